"""world_size-2 gloo worker: GradSync averages every element of a flat gradient arena exactly once, in backward
order, and honours require_backward_grad_sync-style skipping (no collective when not triggered)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_music_generation_b200.ddp import GradSync  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n_layers, per_layer, head, tail = 6, 512, 300, 40
    total = head + n_layers * per_layer + tail
    grad = torch.arange(total, dtype=torch.float32) * (rank + 1)
    layer_ranges = [(head + per_layer * i, head + per_layer * (i + 1)) for i in range(n_layers)]
    sync = GradSync(grad, layer_ranges, (0, head), (head + n_layers * per_layer, total), bucket_mb=1024 * 4 / (1024 * 1024))
    for li in range(n_layers - 1, -1, -1):
        sync.layer_done(li)
    launched_before_tail = list(sync.launched)
    sync.backward_done()
    order = list(sync.launched)
    sync.wait()
    expect = torch.arange(total, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
    assert torch.allclose(grad, expect), (grad - expect).abs().max()
    assert sum(hi - lo for lo, hi in order) == total
    assert all(lo >= head for lo, hi in launched_before_tail)
    assert order[0][1] == head + n_layers * per_layer  # the last block is reduced first
    # deferred final buckets (hierarchical model: the patch embedding's gradient arrives after the stack's backward plan)
    extra = 64
    total2 = total + extra
    grad2 = torch.arange(total2, dtype=torch.float32) * (rank + 1)
    blocks_end = head + n_layers * per_layer
    sync2 = GradSync(grad2, layer_ranges, (0, head), (blocks_end + extra, total2), bucket_mb=1024 * 4 / (1024 * 1024),
                     extra_ranges=[(blocks_end, blocks_end + extra)], defer_final=True)
    for li in range(n_layers - 1, -1, -1):
        sync2.layer_done(li)
    sync2.backward_done()  # must not release the final buckets yet
    assert all(lo >= head and hi <= blocks_end for lo, hi in sync2.launched)
    assert torch.equal(grad2[blocks_end:], torch.arange(blocks_end, total2, dtype=torch.float32) * (rank + 1))
    sync2.finalize()
    assert sum(hi - lo for lo, hi in sync2.launched) == total2
    sync2.wait()
    expect2 = torch.arange(total2, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
    assert torch.allclose(grad2, expect2)
    dist.barrier()
    print("GRADSYNC_OK", rank, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
