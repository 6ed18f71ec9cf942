"""Driver call surface: configurator semantics on CPU; train.py -> ckpt.pt -> sample.py end to end on the GPU with a
synthetic char-level ABC corpus in the reference's on-disk format (uint16 train.bin / val.bin + meta.pkl)."""
import json
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TUNE = "$X:1\nL:1/8\nM:6/8\nK:G\n|: GAB dBG | ABc edc | BAG AFD | G3 G3 :|\n|: gfg edB | dBG ABc | BAG AFD | G3 G3 :|\n"


def test_configurator_overrides(tmp_path):
    from configurator import load_settings
    cfg = tmp_path / "c.py"
    cfg.write_text("batch_size = 8\nout_dir = 'x'\n")
    s = load_settings({"batch_size": 12, "out_dir": "out", "learning_rate": 6e-4, "bias": False},
                      [str(cfg), "--learning_rate=1e-3", "--bias=True"])
    assert s["batch_size"] == 8 and s["out_dir"] == "x" and s["learning_rate"] == 1e-3 and s["bias"] is True
    with pytest.raises(ValueError):
        load_settings({"a": 1}, ["--nope=3"])
    with pytest.raises(TypeError):
        load_settings({"a": 1}, ["--a=hello"])


def test_lr_schedule_matches_reference_formula():
    sys.path.insert(0, ROOT)
    import importlib
    train = importlib.import_module("train")
    s = dict(warmup_iters=100, lr_decay_iters=1000, learning_rate=1e-3, min_lr=1e-4)
    assert train.lr_at(0, s) == pytest.approx(1e-3 * 1 / 101)
    assert train.lr_at(100, s) == pytest.approx(1e-3)
    assert train.lr_at(550, s) == pytest.approx(1e-4 + 0.5 * 9e-4)
    assert train.lr_at(2000, s) == 1e-4


@pytest.mark.gpu
def test_train_then_sample_end_to_end(tmp_path, cuda_device):
    chars = sorted(set(TUNE))
    stoi = {c: i for i, c in enumerate(chars)}
    ids = np.array([stoi[c] for c in TUNE * 400], dtype=np.uint16)
    d = tmp_path / "data" / "irishman"
    d.mkdir(parents=True)
    ids[: int(0.9 * len(ids))].tofile(d / "train.bin")
    ids[int(0.9 * len(ids)):].tofile(d / "val.bin")
    with open(d / "meta.pkl", "wb") as f:
        pickle.dump({"vocab_size": len(chars), "itos": {i: c for c, i in stoi.items()}, "stoi": stoi}, f)
    env = dict(os.environ, PYTHONPATH=ROOT)
    common = ["--dataset=irishman", "--out_dir=out-test"]
    r = subprocess.run([sys.executable, os.path.join(ROOT, "train.py"), *common, "--n_layer=2", "--n_head=2", "--n_embd=128",
                        "--block_size=128", "--batch_size=16", "--gradient_accumulation_steps=2", "--max_iters=60",
                        "--eval_interval=30", "--eval_iters=4", "--log_interval=10", "--learning_rate=3e-3",
                        "--warmup_iters=5", "--lr_decay_iters=60", "--min_lr=3e-4"],
                       cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    recs = [json.loads(l) for l in open(tmp_path / "out-test" / "losses.jsonl")]
    assert [x["step"] for x in recs] == [0, 30, 60]
    assert recs[-1]["val_loss"] < recs[0]["val_loss"] - 1.0       # the repeated tune is learnable in 60 steps
    assert os.path.exists(tmp_path / "out-test" / "ckpt.pt")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "sample.py"), *common, "--tokens_format=char",
                        "--use_validation_prefixes=False", "--num_samples=3", "--max_new_tokens=40", "--top_k=1"],
                       cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    outs = sorted(os.listdir(tmp_path / "out-test" / "samples"))
    assert outs == ["sample_0.abc", "sample_1.abc", "sample_2.abc"]
    texts = [open(tmp_path / "out-test" / "samples" / o).read() for o in outs]
    assert texts[0].startswith("X:0\n") and texts[0][4:] == texts[1][4:]   # greedy: identical continuations


@pytest.mark.gpu
def test_device_token_stream_is_bit_exact_with_reference_get_batch(tmp_path, cuda_device):
    """Integer path: the device gather must reproduce the reference's get_batch (train.py:136-138) bit for bit, for uint16
    and uint32 corpora, given the same torch seed."""
    import torch
    from ai_music_generation_b200 import DeviceTokenStream
    rng = np.random.default_rng(0)
    for wide, dtype in ((False, np.uint16), (True, np.uint32)):
        d = tmp_path / ("w" if wide else "n")
        d.mkdir()
        data = rng.integers(0, 65535 if not wide else 98465, size=50_000).astype(dtype)
        data.tofile(d / "train.bin")
        data[:5000].tofile(d / "val.bin")
        B, T = 16, 256
        stream = DeviceTokenStream(str(d), T, B, cuda_device, wide_tokens=wide)
        for split, arr in (("train", data), ("val", data[:5000])):
            torch.manual_seed(123)
            x, y = stream.get(split)
            torch.manual_seed(123)
            ix = torch.randint(len(arr) - T, (B,))
            xr = torch.stack([torch.from_numpy(arr[i:i + T].astype(np.int64)) for i in ix])
            yr = torch.stack([torch.from_numpy(arr[i + 1:i + 1 + T].astype(np.int64)) for i in ix])
            assert torch.equal(x.cpu(), xr) and torch.equal(y.cpu(), yr)
