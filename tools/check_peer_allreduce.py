#!/usr/bin/env python3
"""torchrun --nproc-per-node N tools/check_peer_allreduce.py
The peer-memory all-reduce kernel (csrc/peer_allreduce.cu) against torch.distributed's NCCL all_reduce:
odd sizes, unaligned views, many tensors per launch, repeated calls (epoch / parity logic), CUDA-graph replay,
and bit-identical results on every rank."""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as td


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/null")
    from timegan_b200 import dist as D
    D.init(backend="nccl", device=dev)
    pc = D.peer_comm()
    assert pc is not None, "peer comm not enabled"
    ok = True
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    # the last one is a 24 MB bucket (the H = 256 model's gradients): 1 465 chunks, more than the kernel's grid cap, so
    # every CTA walks several chunks
    sizes = [1, 3, 4, 17, 4096, 4097, 12288, 100001, 290830, 6000000]
    for rep in range(6):
        D.begin_step("T")
        base = [torch.randn(n + 1, generator=g, device=dev) for n in sizes]
        tens = [b[1:] if i % 2 else b[:-1] for i, b in enumerate(base)]      # odd ones are 4-byte-offset views
        tens = [t if t.is_contiguous() else t.contiguous() for t in tens]
        ref = [t.clone() for t in tens]
        for r in ref:
            td.all_reduce(r)
        pc.allreduce_(tens)
        torch.cuda.synchronize()
        for t, r, n in zip(tens, ref, sizes):
            err = (t - r).abs().max().item()
            scale = r.abs().max().item() + 1e-6
            if err > 1e-5 * scale:
                ok = False
                print(f"rank {rank} rep {rep} size {n}: max err {err:.3e}", flush=True)
        # bit-identical across ranks
        chk = torch.cat([t.view(torch.int32).to(torch.int64).sum().reshape(1) for t in tens])
        lo, hi = chk.clone(), chk.clone()
        td.all_reduce(lo, op=td.ReduceOp.MIN); td.all_reduce(hi, op=td.ReduceOp.MAX)
        if not torch.equal(lo, hi):
            ok = False
            print(f"rank {rank} rep {rep}: results differ between ranks", flush=True)

    # graph replay: x <- allreduce(x * 0.5 + rank_const) for 20 replays vs the same thing eagerly with NCCL
    x = torch.arange(50000, device=dev, dtype=torch.float32) / 50000 + rank
    y = x.clone()
    side = torch.cuda.Stream()
    def body(v):
        v.mul_(0.5).add_(float(rank + 1))
        D.begin_step("G")
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            pc.allreduce_([v])
        torch.cuda.current_stream().wait_stream(side)
        v.mul_(1.0 / world)
    for _ in range(2):
        body(x)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, capture_error_mode="thread_local"):
        body(x)
    for _ in range(20):
        gr.replay()
    for _ in range(23):
        y.mul_(0.5).add_(float(rank + 1))
        td.all_reduce(y)
        y.mul_(1.0 / world)
    torch.cuda.synchronize()
    err = (x - y).abs().max().item()
    if err > 1e-4:
        ok = False
        print(f"rank {rank}: graph replay deviates from NCCL by {err:.3e}", flush=True)

    # latency of one bucket the size of c2's G+S+E+R gradients
    big = [torch.randn(290830, device=dev)]
    D.begin_step("L")
    for _ in range(5):
        D.begin_step("L"); pc.allreduce_(big)
    td.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        D.begin_step("L"); pc.allreduce_(big)
    e1.record(); torch.cuda.synchronize()
    t_peer = e0.elapsed_time(e1) / 50
    e0.record()
    for _ in range(50):
        td.all_reduce(big[0])
    e1.record(); torch.cuda.synchronize()
    t_nccl = e0.elapsed_time(e1) / 50
    pc.check_status()
    print(f"rank {rank}: {'OK' if ok else 'MISMATCH'}  1.16 MB bucket: peer kernel {t_peer*1e3:.1f} us, NCCL {t_nccl*1e3:.1f} us",
          flush=True)
    D.shutdown()
    td.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
