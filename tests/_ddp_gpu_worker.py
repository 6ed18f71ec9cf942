"""torchrun worker (one process per GPU, NCCL): DDP gradient == single-process gradient of the concatenated batch
(with equal per-rank batches and no ignored targets W-rank DDP equals the concat-batch gradient), the
accumulation toggle skips the collective on non-final micro-steps, and ranks stay bitwise in sync after optimizer steps."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_music_generation_b200 import DDP, GPT, GPTConfig  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = dict(block_size=128, vocab_size=95, n_layer=4, n_head=2, n_embd=128, dropout=0.0, bias=False)
    torch.manual_seed(1337 + rank)  # different init per rank: the DDP constructor must broadcast rank 0's parameters
    model = GPT(GPTConfig(**cfg)).to(dev).train()
    ddp = DDP(model, bucket_mb=0.5)
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
    norm = model.clip_grad_norm_(1.0)
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
    print(f"DDP_GPU_OK rank {rank} rel {rel:.2e} norm {norm.item():.4f} buckets {len(model._grad_sync.buckets)}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
