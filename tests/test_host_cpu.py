"""CPU-side tests: the C-ABI library loads and exports every declared symbol, host logic of the module / optimizer /
DDP bucket planner, and the multi-rank gradient averaging over gloo (world_size 2)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from ai_music_generation_b200 import _C, build
    build.build()
    lib = _C.lib()
    declared = _C.declared_symbols()
    assert len(declared) >= 17
    for sym in declared:
        assert hasattr(lib, sym), sym
        assert sym in _C._SIGNATURES, f"{sym} declared in include/abcgpt.h but not bound"
    assert lib.abcgpt_version() == 100


def test_state_dict_and_init_match_reference_layout():
    from ai_music_generation_b200 import GPT, GPTConfig
    from oracle import nanogpt_oracle as O
    cfgd = dict(block_size=32, vocab_size=95, n_layer=2, n_head=2, n_embd=128, dropout=0.0, bias=True)
    torch.manual_seed(1337)
    m = GPT(GPTConfig(**cfgd))
    names = O.param_names(O.OracleConfig(**cfgd))
    sd = m.state_dict()
    assert list(sd.keys()) == names + ["lm_head.weight"]
    assert sd["lm_head.weight"].data_ptr() == sd["transformer.wte.weight"].data_ptr()
    shapes = O.param_shapes(O.OracleConfig(**cfgd))
    for n in names:
        assert tuple(sd[n].shape) == shapes[n], n
    # c_proj scaled init (model.py:143-145)
    assert sd["transformer.h.0.mlp.c_proj.weight"].std().item() == pytest.approx(0.02 / 2.0, rel=0.1)
    # arena: decay tensors first, 16-byte aligned starts, parameters are views of one buffer
    a = m._arena
    flat = a["flat"]
    for p, o in zip(a["params"], a["offs"]):
        assert o % 8 == 0
        assert p.data_ptr() == flat.data_ptr() + 4 * o
    dims = [p.dim() >= 2 for p in a["params"]]
    assert dims == sorted(dims, reverse=True)
    # round trip through load_state_dict keeps the views
    sd2 = {k: v.clone() + 1.0 for k, v in sd.items()}
    m.load_state_dict(sd2)
    assert torch.equal(m.state_dict()["transformer.wpe.weight"], sd2["transformer.wpe.weight"])
    assert a["params"][0].data_ptr() == flat.data_ptr()


def test_configure_optimizers_groups_and_state_dict():
    from ai_music_generation_b200 import GPT, GPTConfig
    m = GPT(GPTConfig(block_size=32, vocab_size=95, n_layer=6, n_head=6, n_embd=384, dropout=0.0, bias=False))
    opt = m.configure_optimizers(0.1, 1e-3, (0.9, 0.99), "cuda")
    g0, g1 = opt.param_groups
    assert len(g0["params"]) == 26 and len(g1["params"]) == 13
    assert sum(p.numel() for p in g0["params"]) == 10_751_616 - (256 - 32) * 384  # SURVEY: 10 751 616 at block 256
    assert g0["weight_decay"] == 0.1 and g1["weight_decay"] == 0.0
    for g in opt.param_groups:
        g["lr"] = 5e-4                                                   # train.py:285-287
    sd = opt.state_dict()
    assert len(sd["param_groups"]) == 2 and sd["param_groups"][0]["lr"] == 5e-4
    opt.zero_grad(set_to_none=True)
    assert opt.step() is None                                            # no grads: no-op, no GPU needed


def test_forward_refuses_cpu_tensors():
    from ai_music_generation_b200 import GPT, GPTConfig, _C
    m = GPT(GPTConfig(block_size=16, vocab_size=95, n_layer=1, n_head=1, n_embd=64, dropout=0.0, bias=False))
    with pytest.raises(_C.AbcgptError):
        m(torch.zeros(1, 8, dtype=torch.long), torch.zeros(1, 8, dtype=torch.long))


def test_bucket_plan_covers_arena_once():
    from ai_music_generation_b200.ddp import plan_buckets
    layer_ranges = [(1000 + 700 * i, 1000 + 700 * (i + 1)) for i in range(12)]
    buckets = plan_buckets(layer_ranges, (0, 1000), (9400, 9500), bucket_elems=2000)
    covered = sorted((lo, hi) for _, lo, hi in buckets)
    assert covered[0][0] == 0 and covered[-1][1] == 9500
    for (a, b), (c, d) in zip(covered, covered[1:]):
        assert b == c
    # launch order follows the backward: last layers first, embedding / 1-D tail last
    triggers = [t for t, _, _ in buckets]
    assert triggers[:-2] == sorted(triggers[:-2], reverse=True) and triggers[-2:] == [-1, -1]
    assert all(hi - lo >= 2000 for t, lo, hi in buckets[:-3])


def test_bucket_plan_extra_range_after_the_last_block():
    """Hierarchical model: the patch embedding sits between the last block and the 1-D tail of the patch-level arena."""
    from ai_music_generation_b200.ddp import plan_buckets
    layer_ranges = [(1000 + 700 * i, 1000 + 700 * (i + 1)) for i in range(3)]
    buckets = plan_buckets(layer_ranges, (0, 1000), (3600, 3700), bucket_elems=100, extra_ranges=[(3100, 3600)])
    covered = sorted((lo, hi) for _, lo, hi in buckets)
    assert covered[0][0] == 0 and covered[-1][1] == 3700
    for (a, b), (c, d) in zip(covered, covered[1:]):
        assert b == c
    assert [(t, lo, hi) for t, lo, hi in buckets if t == -1] == [(-1, 0, 1000), (-1, 3100, 3600), (-1, 3600, 3700)]


def test_nvls_final_ranges_are_merged_and_still_tile_the_arena():
    """NVLS exchange (ddp.merge_final_buckets): the lowest layer group joins the ranges released at the end of the backward and
    adjacent ranges become one launch; the set of exchanged elements must not change."""
    from ai_music_generation_b200.ddp import merge_final_buckets, plan_buckets
    layer_ranges = [(1000 + 700 * i, 1000 + 700 * (i + 1)) for i in range(12)]
    for extra in ((), [(9400, 9500)]):
        tail = (9500, 9560) if extra else (9400, 9460)
        buckets = plan_buckets(layer_ranges, (0, 1000), tail, bucket_elems=700, extra_ranges=extra)
        merged = merge_final_buckets(buckets)
        covered = sorted((lo, hi) for _, lo, hi in merged)
        assert covered[0][0] == 0 and covered[-1][1] == tail[1]
        for (a, b), (c, d) in zip(covered, covered[1:]):
            assert b == c
        final = [(lo, hi) for t, lo, hi in merged if t == -1]
        assert final == [(0, 1700), (9400, tail[1])]          # head | layer 0, and (patch embedding |) 1-D tail
        assert [t for t, _, _ in merged if t > 0] == list(range(11, 0, -1))   # the other layers keep their release points
        assert len(merged) == len(buckets) - (2 if extra else 1)


def test_dropout_key_patches_of_a_recorded_plan():
    """Launch plans recorded with dropout on are replayed with the next step's keys (ops.key_patches): every launch whose C
    signature ends (..., dropout_p, dropout_key, stream) with p > 0 is a site; identical keys on two sites make the plan unusable."""
    from ai_music_generation_b200 import ops
    keys = [11, 22, 33, 44]
    plan = [("embed_fwd", 1, None, (1, 2, 3, 0.2, 11, 99), ()), ("gemm", 1, None, (5, 6, 0.0, 0, 99), ()),
            ("gemm", 1, None, (5, 6, 0.2, 33, 99), ()), ("layernorm_bwd", 1, None, (7, 0.2, 44, 99), ()), ("py", lambda: None),
            ("sumsq", 2, None, (1, 2, 3), ())]
    plan = [e if e[0] == "py" else e for e in plan]
    patches = ops.key_patches(plan, keys)
    assert patches == [(0, 0), (2, 2), (3, 3)]
    new_keys = [101, 202, 303, 404]
    for i, site in patches:
        plan[i][3][-2] = new_keys[site]
    assert plan[0][3][-2] == 101 and plan[2][3][-2] == 303 and plan[3][3][-2] == 404 and plan[1][3][-2] == 0
    assert ops.key_patches([("gemm", 1, None, (0.2, 5, 99), ())], [5, 5]) is None        # two sites share a key
    assert ops.key_patches([("gemm", 1, None, (0.2, 77, 99), ())], [5, 6]) is None       # a key that is not one of the step's


def test_gradsync_two_ranks_gloo():
    script = os.path.join(ROOT, "tests", "_gloo_gradsync_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29531", script],
                         capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count("GRADSYNC_OK") == 2


def test_affine_plan_deltas_and_program_encoding():
    """Host logic of the decode replay (ops.plan_deltas / ops.compile_affine): arguments that differ between two recorded steps
    must be integers with a constant difference; the packed program interleaves (value, delta) per argument."""
    import struct
    from ai_music_generation_b200 import _C, ops

    def fake(name):
        def f(*a):
            return 0
        f.__name__ = name
        return f

    g, ln = fake("abcgpt_gemm_bf16"), fake("abcgpt_layernorm_fwd")
    plan = lambda t: [("layernorm_fwd", 1, ln, (1000, 2000, 0, 3000, 0, 4000, 5000, 256, 768, 7), (256, 768)),  # noqa: E731
                      ("gemm", 1, g, (10, 0, 768, 20, 0, 768, 256, 2304, 768, 0, 9000 + 4608 * t, 2304 * 1024, 0, 0, 0, 0, 0, 128, 0,
                                      0.0, 0, 7), (256, 2304, 768, 0, 0, 0))]
    d01, d12 = ops.plan_deltas(plan(0), plan(1)), ops.plan_deltas(plan(1), plan(2))
    assert d01 == d12 == [[], [(10, 4608)]]
    assert ops.plan_deltas(plan(0), plan(0)[:1]) is None                       # different structure
    bad = plan(1)
    bad[1] = bad[1][:3] + (bad[1][3][:19] + (0.5,) + bad[1][3][20:],) + bad[1][4:]
    assert ops.plan_deltas(plan(0), bad) is None                               # a float argument may not vary
    arr, n, launches = ops.compile_affine(plan(2), d12)
    words = list(arr)
    assert n == len(words) == (2 + 2 * 10) + (2 + 2 * 22) and launches == 2
    assert words[:2] == [_C.FN_IDS["abcgpt_layernorm_fwd"], 10] and words[2:6] == [1000, 0, 2000, 0]
    gemm = words[22:]
    assert gemm[:2] == [_C.FN_IDS["abcgpt_gemm_bf16"], 22]
    assert gemm[2 + 2 * 10: 4 + 2 * 10] == [9000 + 4608 * 2, 4608]              # the cache row pointer advances per position
    assert gemm[2 + 2 * 19] == struct.unpack("<q", struct.pack("<d", 0.0))[0]    # floats travel as double bit patterns
    assert ops.compile_affine([("x", 1, fake("abcgpt_unknown"), (1,), None)], [[]]) is None


def test_dropout_numpy_twin_matches_the_kernel_header(tmp_path):
    """csrc/dropout.cuh is __host__ __device__: compiled here with g++ (tests/helpers/dropout_host.cpp) it must produce exactly
    the masks ai_music_generation_b200/dropout.py regenerates for the oracle — both generators (murmur lanes for the residual /
    embedding sites, Weyl + folded multiply with 15-bit lanes and the carry-trick comparison for the attention site)."""
    import shutil
    import numpy as np
    from ai_music_generation_b200 import dropout as D
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path / "dropout_host")
    src = os.path.join(ROOT, "tests", "helpers", "dropout_host.cpp")
    r = subprocess.run([gxx, "-O1", "-std=c++17", "-I", os.path.join(ROOT, "ai_music_generation_b200", "csrc"), src, "-o", exe],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    for key, rows, cols, p in ((12345, 37, 70, 0.2), (0xDEADBEEF, 64, 129, 0.1), (7, 5, 33, 0.5)):
        out = subprocess.run([exe, str(key), str(rows), str(cols), str(p)], capture_output=True, text=True, timeout=60)
        assert out.returncode == 0
        resid_s, attn_s = out.stdout.split("\n")[:2]
        resid = np.frombuffer(resid_s.encode(), dtype=np.uint8).reshape(rows, cols) == ord("1")
        attn = np.frombuffer(attn_s.encode(), dtype=np.uint8).reshape(rows, cols) == ord("1")
        assert np.array_equal(resid, D.keep_mask(key, rows, cols, p))
        # attention site: rows = (b*H + h)*T + q with B = H = 1, T = cols needs rows <= T; use the row counter directly
        T = cols
        host = D.attention_keep_mask(key, 1, (rows + T - 1) // T + 1, T, p).reshape(-1, T)[:rows]
        assert np.array_equal(attn, host)
        tol = 4.0 * (p * (1 - p) / (rows * cols)) ** 0.5
        assert abs(attn.mean() - (1 - p)) < tol and abs(resid.mean() - (1 - p)) < tol


def test_dynamic_tile_scheduler_ticket_protocol_model():
    """Model of the ticket protocol of gemm2_kernel's opt-in dynamic scheduler (csrc/gemm.cu): every pair processes its static
    first item, then draws tickets (one held in flight); tickets < dyn_work map to item num_pairs + ticket, the first ticket
    >= dyn_work ends the pair, which never draws again; the holder of ticket dyn_work + num_pairs - 1 resets the counter.
    Under arbitrary interleavings (including pairs that start after all work is gone) every item is processed exactly
    once, every pair draws exactly one end ticket and the counter is back at zero for the next launch."""
    import random
    rng = random.Random(0)
    for trial in range(300):
        num_pairs = rng.randint(1, 9)
        total_work = rng.randint(num_pairs + 1, 60)
        dyn_work = total_work - num_pairs
        counter = [0]
        resets = [0]
        done_items = []
        end_tickets = [0] * num_pairs
        # per pair: pc 0 = not started; state = (pending ticket or None, drew_end)
        pend = [None] * num_pairs
        drew_end = [False] * num_pairs
        phase = ["start"] * num_pairs   # start -> (draw pending, process static) -> loop: publish(pending), draw -> ...
        active = list(range(num_pairs))
        while active:
            i = rng.choice(active)
            if phase[i] == "start":
                pend[i] = counter[0]; counter[0] += 1           # pending = draw()
                done_items.append(i)                             # static first item
                phase[i] = "loop"
                continue
            # end of a tile: publish(it + 1, pending); pending = draw()
            w = pend[i]
            nxt = None
            if w is not None:
                if w == dyn_work + num_pairs - 1:
                    counter[0] = 0; resets[0] += 1
                if w >= dyn_work:
                    drew_end[i] = True; end_tickets[i] += 1
                else:
                    nxt = num_pairs + w
            if not drew_end[i]:
                pend[i] = counter[0]; counter[0] += 1
            else:
                pend[i] = None
            if nxt is None:
                active.remove(i)                                 # published -1: every role of this pair breaks
            else:
                done_items.append(nxt)
        assert sorted(done_items) == list(range(total_work)), (trial, num_pairs, total_work)
        assert end_tickets == [1] * num_pairs and resets[0] == 1 and counter[0] == 0


def test_from_pretrained_weight_mapping_matches_the_reference_rules():
    """GPT.load_hf_gpt2_state_dict (the body of from_pretrained, nanoGPT/model.py:236-259) on a randomly initialised HF GPT-2
    of a small shape: Conv1D weights transposed, mask buffers dropped, tied head, everything else copied verbatim.  (The
    hub download itself needs network access and is not exercised.)"""
    transformers = pytest.importorskip("transformers")
    from ai_music_generation_b200 import GPT, GPTConfig
    torch.manual_seed(0)
    hf = transformers.GPT2LMHeadModel(transformers.GPT2Config(n_layer=2, n_head=2, n_embd=128, vocab_size=300, n_positions=64))
    sd_hf = hf.state_dict()
    m = GPT(GPTConfig(block_size=64, vocab_size=300, n_layer=2, n_head=2, n_embd=128, dropout=0.0, bias=True))
    m.load_hf_gpt2_state_dict(sd_hf)
    sd = m.state_dict()
    assert torch.equal(sd["transformer.h.1.mlp.c_fc.weight"], sd_hf["transformer.h.1.mlp.c_fc.weight"].t())
    assert torch.equal(sd["transformer.h.0.attn.c_attn.weight"], sd_hf["transformer.h.0.attn.c_attn.weight"].t())
    assert torch.equal(sd["transformer.h.0.attn.c_attn.bias"], sd_hf["transformer.h.0.attn.c_attn.bias"])
    assert torch.equal(sd["transformer.wpe.weight"], sd_hf["transformer.wpe.weight"])
    assert torch.equal(sd["lm_head.weight"], sd_hf["transformer.wte.weight"]) and torch.equal(sd["transformer.ln_f.weight"], sd_hf["transformer.ln_f.weight"])
    with pytest.raises(ValueError):
        GPT.from_pretrained("gpt3")
    with pytest.raises(ValueError):
        GPT.from_pretrained("gpt2", dict(bias=False))
    bad = dict(sd_hf)
    bad.pop("transformer.ln_f.bias")
    with pytest.raises(ValueError):
        m.load_hf_gpt2_state_dict(bad)


def test_patchilizer_matches_the_reference_codec_and_sampling_helpers():
    """Bar <-> patch codec against vectors produced by the UNMODIFIED reference Patchilizer (tunesformer/utils.py:9-82,
    oracle/make_golden_tunesformer.py), and the restated sampling helpers' invariants."""
    import json
    import os
    import numpy as np
    from ai_music_generation_b200.tunesformer import Patchilizer, temperature_draw, top_k_filter, top_p_filter
    with open(os.path.join(os.path.dirname(__file__), "golden", "tunesformer_tiny_generate.json")) as f:
        g = json.load(f)
    pz = Patchilizer()
    for c in g["codec"]:
        assert pz.encode(c["text"]) == c["plain"], c["text"]
        assert pz.encode(c["text"], add_special_patches=True) == c["special"], c["text"]
        assert pz.decode(c["special"]) == c["decoded"]
    assert pz.bar2patch("x" * 100) == [1] + [ord("x")] * 31 and pz.patch2bar([1, 65, 66, 2, 0, 0]) == "AB"
    p = np.array([0.05, 0.4, 0.3, 0.15, 0.1])
    assert np.allclose(top_p_filter(p, 0.6), [0, 4 / 7, 3 / 7, 0, 0]) and np.allclose(top_p_filter(p, 1.0), p)
    assert np.allclose(top_k_filter(p, 2), [0, 4 / 7, 3 / 7, 0, 0]) and np.allclose(top_k_filter(p, 0), p)
    assert temperature_draw(top_k_filter(p, 1), 1.2, seed=3) == 1 and temperature_draw(p, 0.0) == 1
    draws = [temperature_draw(p, 1.0, seed=s) for s in range(400)]
    assert abs(np.mean(np.array(draws) == 1) - 0.4) < 0.08 and temperature_draw(p, 1.0, seed=7) == temperature_draw(p, 1.0, seed=7)
