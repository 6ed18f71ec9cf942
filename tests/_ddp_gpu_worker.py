"""torchrun worker (one process per GPU, NCCL): DDP gradient == single-process gradient of the concatenated batch
(with equal per-rank batches and no ignored targets W-rank DDP equals the concat-batch gradient), the
accumulation toggle skips the collective on non-final micro-steps, and ranks stay bitwise in sync after optimizer steps."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_music_generation_b200 import DDP, GPT, GPTConfig, TunesFormerDDP, TunesFormerShaped  # noqa: E402


def tunesformer_case(rank, world, dev):
    """Hierarchical model (BASELINE config 4) under TunesFormerDDP: both arenas' gradients == those of the concatenated
    batch (every rank's shard has the same pad pattern, so the mean of per-rank mean losses is the global mean), the patch
    embedding's late gradient is part of the exchange, replicas stay bitwise in sync over optimizer steps."""
    pc = dict(block_size=16, vocab_size=128, n_layer=2, n_head=2, n_embd=128, dropout=0.0, bias=True, activation="gelu_tanh")
    cc = dict(block_size=32, vocab_size=128, n_layer=2, n_head=2, n_embd=128, dropout=0.0, bias=True, activation="gelu_tanh")
    torch.manual_seed(99 + rank)
    model = TunesFormerShaped(GPTConfig(**pc), GPTConfig(**cc)).to(dev).train()
    ddp = TunesFormerDDP(model, bucket_mb=0.25)
    ref = TunesFormerShaped(GPTConfig(**pc), GPTConfig(**cc)).to(dev).train()
    ref.patch_level_decoder.load_state_dict(model.patch_level_decoder.state_dict())
    ref.char_level_decoder.load_state_dict(model.char_level_decoder.state_dict())
    g = torch.Generator().manual_seed(11)
    Bt = 3
    lens = torch.randint(6, 33, (Bt, 12), generator=g)
    shards = []
    for _ in range(world):
        pt = torch.randint(3, 128, (Bt, 12, 32), generator=g)
        pt[torch.arange(32)[None, None, :] >= lens[..., None]] = 0
        shards.append(pt)
    full = torch.cat(shards, 0).to(dev)
    opt = model.configure_optimizers(0.1, 1e-3, (0.9, 0.95))
    loss = ddp(shards[rank].to(dev))
    loss.backward()
    for dec in (model.patch_level_decoder, model.char_level_decoder):
        dec._grad_sync.wait()
    torch.cuda.synchronize()
    ref(full).backward()
    rels = []
    for dec, rdec in ((model.patch_level_decoder, ref.patch_level_decoder), (model.char_level_decoder, ref.char_level_decoder)):
        got, want = dec._arena["grad"], rdec._arena["grad"]
        rels.append(((got - want).norm() / want.norm()).item())
        assert rels[-1] < 2e-2, rels
        pe = [n for n in dec._arena["names"] if n.startswith("patch_embedding")]
        for n in pe:   # the late gradient took part in the exchange
            a, b = dec._view("grad", n).float(), rdec._view("grad", n).float()
            assert ((a - b).norm() / b.norm()).item() < 2e-2, n
    for _ in range(3):
        model.clip_grad_norm_(1.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
        ddp(shards[rank].to(dev)).backward()
    for dec in (model.patch_level_decoder, model.char_level_decoder):
        dec._grad_sync.wait()
        flat = dec._arena["flat"]
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        for other in gathered:
            assert torch.equal(other, flat), "hierarchical replicas diverged"
    print(f"DDP_TUNESFORMER_OK rank {rank} rel {rels[0]:.2e} {rels[1]:.2e} buckets "
          f"{len(model.patch_level_decoder._grad_sync.buckets)}+{len(model.char_level_decoder._grad_sync.buckets)}", flush=True)


def stock_train_loop_case(rank, world, dev):
    """nanoGPT/train.py:335-357 verbatim (GradScaler(enabled=False), the STOCK torch.nn.utils.clip_grad_norm_ over
    model.parameters(), scaler.step) with our DDP class in place of torch's: no explicit wait anywhere in the loop — the
    backward itself must hand back fully reduced gradients (GradSync.finalize joins the compute stream).  Held to a
    single-process run of the same loop over the concatenated batch, and replicas must stay bitwise identical."""
    from ai_music_generation_b200 import evalloop
    cfg = dict(block_size=128, vocab_size=95, n_layer=4, n_head=2, n_embd=128, dropout=0.0, bias=False)
    torch.manual_seed(4242)   # same init everywhere
    model = GPT(GPTConfig(**cfg)).to(dev).train()
    ref = GPT(GPTConfig(**cfg)).to(dev).train()
    ref.load_state_dict(model.state_dict())
    raw_model = model
    model = DDP(model, device_ids=[dev.index], bucket_mb=0.25, nvls=True)   # NVLS exchange where the box has multicast memory (else NCCL)
    g = torch.Generator().manual_seed(21)
    B, accum, iters, grad_clip = 2, 2, 4, 1.0
    data = [(torch.randint(95, (world * B, 128), generator=g), torch.randint(95, (world * B, 128), generator=g))
            for _ in range(accum * iters + 1)]
    ctx = torch.amp.autocast(device_type="cuda", dtype=torch.bfloat16)
    results = []
    for net, sl, is_ddp in ((model, slice(rank * B, (rank + 1) * B), True), (ref, slice(None), False)):
        raw = raw_model if is_ddp else ref
        scaler = torch.amp.GradScaler("cuda", enabled=False)
        optimizer = raw.configure_optimizers(0.1, 1e-3, (0.9, 0.95), "cuda")
        it = iter(data)

        def get_batch(split):
            x, y = next(it)
            return x[sl].contiguous().to(dev), y[sl].contiguous().to(dev)

        ddp, gradient_accumulation_steps = is_ddp, accum
        norms = []
        X, Y = get_batch("train")
        for iter_num in range(iters):
            for micro_step in range(gradient_accumulation_steps):
                if ddp:
                    net.require_backward_grad_sync = micro_step == gradient_accumulation_steps - 1
                with ctx:
                    logits, loss = net(X, Y)
                    loss = loss / gradient_accumulation_steps
                X, Y = get_batch("train")
                scaler.scale(loss).backward()
            if grad_clip != 0.0:
                scaler.unscale_(optimizer)
                norms.append(torch.nn.utils.clip_grad_norm_(net.parameters(), grad_clip).item())
            scaler.step(optimizer)
            scaler.update()
            optimizer.zero_grad(set_to_none=True)
        results.append((norms, raw._arena["flat"].clone()))
    (n_ddp, p_ddp), (n_ref, p_ref) = results
    for a, b in zip(n_ddp, n_ref):
        assert abs(a - b) <= 2e-2 * b, (n_ddp, n_ref)
    rel = ((p_ddp - p_ref).norm() / p_ref.norm()).item()
    assert rel < 2e-3, rel
    gathered = [torch.empty_like(p_ddp) for _ in range(world)]
    dist.all_gather(gathered, p_ddp)
    for other in gathered:
        assert torch.equal(other, p_ddp), "replicas diverged under the stock loop"
    # rank-sharded estimate_loss (train.py:231-244 / SURVEY 8f N4): world x ceil(n / world) batches, one all-reduce per split
    evb = [(torch.randint(95, (2, 128), generator=g), torch.randint(95, (2, 128), generator=g)) for _ in range(4)]
    cur = [0]

    def eval_batch(split):   # rank r takes batches r, r + world, ...
        x, y = evb[(rank + world * cur[0]) % 4]
        cur[0] += 1
        return x.to(dev), y.to(dev)

    out = evalloop.estimate_loss(raw_model, eval_batch, 4, dev, shard=True, splits=("val",))
    raw_model.eval()
    with torch.no_grad():
        want = sum(raw_model(x.to(dev), y.to(dev))[1].item() for x, y in evb) / 4
    raw_model.train()
    if world in (1, 2, 4):
        assert abs(out["val"].item() - want) <= 1e-5, (out["val"].item(), want)
    print(f"DDP_STOCK_LOOP_OK rank {rank} rel {rel:.2e} norms {n_ddp[0]:.4f}/{n_ref[0]:.4f} eval {out['val'].item():.5f}", flush=True)


def gpt_case(rank, world, dev, nvls):
    """DDP gradient == concatenated-batch gradient, the accumulation toggle, replicas bitwise in sync after optimizer steps.
    nvls=True: the fused NVLink-switch exchange (csrc/nvls.cu) instead of NCCL buckets — same checks, plus the norm that
    comes with the gradients must be the norm of the arena."""
    cfg = dict(block_size=128, vocab_size=95, n_layer=4, n_head=2, n_embd=128, dropout=0.0, bias=False)
    torch.manual_seed(1337 + rank)  # different init per rank: the DDP constructor must broadcast rank 0's parameters
    model = GPT(GPTConfig(**cfg)).to(dev).train()
    ddp = DDP(model, bucket_mb=0.5, nvls=nvls)
    ref = GPT(GPTConfig(**cfg)).to(dev).train()
    ref.load_state_dict(model.state_dict())
    p0 = model._arena["flat"].clone()
    dist.broadcast(p0, src=0)
    assert torch.equal(p0, model._arena["flat"]), "constructor broadcast failed"

    g = torch.Generator().manual_seed(7)
    B = 4
    X = torch.randint(95, (world * B, 128), generator=g).to(dev)
    Y = torch.randint(95, (world * B, 128), generator=g).to(dev)
    opt = model.configure_optimizers(0.1, 1e-3, (0.9, 0.95), "cuda")
    # two micro-steps: the first must NOT all-reduce
    ddp.require_backward_grad_sync = False
    _, l0 = ddp(X[rank * B:(rank + 1) * B].contiguous(), Y[rank * B:(rank + 1) * B].contiguous())
    if nvls and model._grad_sync.nvls is None:
        return False   # no multicast on this box: the wrapper fell back to NCCL (covered by the nvls=False run)
    (l0 / 2).backward()
    local_only = model._arena["grad"].clone()
    ddp.require_backward_grad_sync = True
    _, l1 = ddp(X[rank * B:(rank + 1) * B].contiguous(), Y[rank * B:(rank + 1) * B].contiguous())
    (l1 / 2).backward()
    model._grad_sync.wait()
    torch.cuda.synchronize()
    got = model._arena["grad"].clone()
    # reference: every rank computes the concatenated batch locally, no communication
    _, lr = ref(X, Y)
    lr.backward()
    want = ref._arena["grad"]
    rel = ((got - want).norm() / want.norm()).item()
    assert rel < 2e-2, rel
    assert not torch.allclose(local_only * 2, got, rtol=1e-3, atol=1e-6), "first micro-step seems to have been all-reduced"
    gathered = [torch.empty_like(got) for _ in range(world)]
    dist.all_gather(gathered, got)
    for other in gathered:
        assert torch.equal(other, got), "exchanged gradients differ between ranks"
    norm = model.clip_grad_norm_(1.0)
    if nvls:
        assert model._grad_sync.norm_fresh, "the fused exchange did not leave a norm"
        assert abs(norm.item() - got.double().norm().item()) <= 1e-4 * got.double().norm().item(), (norm.item(), got.norm().item())
    opt.step()
    opt.zero_grad(set_to_none=True)
    for _ in range(3):
        _, loss = ddp(X[rank * B:(rank + 1) * B].contiguous(), Y[rank * B:(rank + 1) * B].contiguous())
        loss.backward()
        model.clip_grad_norm_(1.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
    flat = model._arena["flat"]
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    for other in gathered:
        assert torch.equal(other, flat), "ranks diverged"
    tag = "DDP_NVLS_OK" if nvls else "DDP_GPU_OK"
    print(f"{tag} rank {rank} rel {rel:.2e} norm {norm.item():.4f} buckets {len(model._grad_sync.buckets)}", flush=True)
    return True


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    gpt_case(rank, world, dev, nvls=False)
    if not gpt_case(rank, world, dev, nvls=True):
        print(f"DDP_NVLS_SKIPPED rank {rank} (no multicast memory on this box)", flush=True)
    tunesformer_case(rank, world, dev)
    stock_train_loop_case(rank, world, dev)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
