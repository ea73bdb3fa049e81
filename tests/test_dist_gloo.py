"""Data-parallel host logic on CPU: two `gloo` ranks (torch.multiprocessing) exercise timegan_b200.dist --
statistics all-reduce, global means (forward and backward), bucketed gradient all-reduce, batch sharding -- and
check the DP convention of SURVEY.md section 8e: every rank evaluates the GLOBAL-batch loss and the summed local
gradient contributions equal the single-process gradient.  The kernels themselves are exercised by the -m gpu
tests; nothing here launches one."""
import os
import socket

import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    td.init_process_group("gloo", rank=rank, world_size=world)
    from timegan_b200 import dist as D
    D.enable()
    try:
        assert D.is_enabled() and D.world_size() == world and D.rank() == rank
        g = torch.Generator().manual_seed(0)
        X = torch.rand(8, 5, 3, generator=g)                 # the global batch, identical on all ranks
        w = torch.rand(3, generator=g).requires_grad_(True)  # a "parameter"
        xs = D.shard_batch(X)
        assert xs.shape[0] == 8 // world and torch.equal(xs, X[rank * 4:(rank + 1) * 4])
        # --- ragged global batch (drop_last=False, tt:33-37): every sequence is used, shards differ by one, and the
        #     statistics stay exact because they are count-weighted sums over the GLOBAL count ---
        xr = D.shard_batch(X[:7])
        assert xr.shape[0] == (4 if rank == 0 else 3) and torch.equal(xr, X[:7][0:4] if rank == 0 else X[:7][4:7])
        assert D.global_count(xr.shape[0]) == 7.0 and D.global_count(xr.numel()) == float(X[:7].numel())
        s7, n7 = D.allreduce_stats(xr.sum((0, 1)), xr.shape[0] * xr.shape[1])
        assert n7 == 35.0 and torch.allclose(s7, X[:7].sum((0, 1)), atol=1e-6)
        wr = torch.rand(3, generator=torch.Generator().manual_seed(5)).requires_grad_(True)
        lm = D.global_mean(((xr * wr).sum(-1) ** 2).sum(), xr.shape[0] * xr.shape[1])
        lm.backward()
        wref = wr.detach().clone().requires_grad_(True)
        ref7 = (((X[:7] * wref).sum(-1)) ** 2).mean()
        ref7.backward()
        assert torch.allclose(lm.detach(), ref7.detach(), atol=1e-6)
        gb = D.GradBuckets(); gb.launch([wr]); gb.wait()     # SUM of the local contributions = the global gradient
        assert torch.allclose(wr.grad, wref.grad, atol=1e-6), (wr.grad, wref.grad)
        assert D.shard_batch(X[:1]) is None                  # fewer sequences than ranks: skipped on every rank alike
        assert D.shard_bounds(9, 8, 0) == (0, 2) and D.shard_bounds(9, 8, 7) == (8, 9)
        xs = D.shard_batch(X)                                # back to the even split for the checks below
        # --- statistics all-reduce: sum + global count ---
        s, n = D.allreduce_stats(xs.sum((0, 1)), xs.shape[0] * xs.shape[1])
        assert n == 40 and torch.allclose(s, X.sum((0, 1)), atol=1e-6)
        # --- a loss that is NOT a mean of per-sample terms: sqrt of the global MSE (recon_loss, tt:72-74) ---
        def local_sse(x):
            return ((x * w).sum(-1) ** 2).sum()
        sse, cnt = D.allreduce_stats(local_sse(xs).detach().reshape(1), xs.shape[0] * xs.shape[1])
        loss_val = torch.sqrt(sse / cnt)
        # local contribution to d loss / d w = (1 / (2 sqrt(.))) * d(local sse / cnt)/dw  (what _ReconLoss.backward does)
        (local_sse(xs) / cnt).backward()
        w.grad.mul_(0.5 / loss_val.item())
        b = D.GradBuckets()
        b.launch([w])
        b.wait()
        w_ref = w.detach().clone().requires_grad_(True)
        ref = torch.sqrt((((X * w_ref).sum(-1)) ** 2).mean())
        ref.backward()
        assert torch.allclose(loss_val, ref.detach().reshape(1), atol=1e-6)
        assert torch.allclose(w.grad, w_ref.grad, atol=1e-6), (w.grad, w_ref.grad)
        # --- differentiable global mean (BCE, accuracy) ---
        p = torch.rand(4, 1, generator=torch.Generator().manual_seed(10 + rank)).requires_grad_(True)
        gm = D.global_mean(p.sum(), p.numel())
        gm.backward()
        allp = [torch.rand(4, 1, generator=torch.Generator().manual_seed(10 + r)) for r in range(world)]
        assert torch.allclose(gm.detach(), torch.cat(allp).mean(), atol=1e-6)
        assert torch.allclose(p.grad, torch.full_like(p, 1.0 / (4 * world)))
        # --- hook-driven bucketed reducer: buckets fire as soon as their last gradient lands ---
        lin1, lin2 = torch.nn.Linear(3, 4), torch.nn.Linear(4, 2)
        with torch.no_grad():
            for q in list(lin1.parameters()) + list(lin2.parameters()):
                q.copy_(torch.rand(q.shape, generator=torch.Generator().manual_seed(int(q.numel()))))
        red = D.GradReducer([list(lin2.parameters()), list(lin1.parameters())])
        red.arm()
        out = lin2(torch.tanh(lin1(xs))).sum() / (40 * 2)          # local contribution to the global mean
        out.backward()
        assert red.launched == [True, True]                          # both fired from the hooks, before finish()
        red.finish()
        ref1, ref2 = torch.nn.Linear(3, 4), torch.nn.Linear(4, 2)
        ref1.load_state_dict(lin1.state_dict()); ref2.load_state_dict(lin2.state_dict())
        (ref2(torch.tanh(ref1(X))).sum() / (40 * 2)).backward()
        for a, b in zip(list(lin1.parameters()) + list(lin2.parameters()), list(ref1.parameters()) + list(ref2.parameters())):
            assert torch.allclose(a.grad, b.grad, atol=1e-6)
        red.remove()
        ret[rank] = "ok"
    finally:
        D.disable()
        td.destroy_process_group()


def test_dp_host_logic_two_gloo_ranks():
    pytest.importorskip("timegan_b200")
    world = 2
    port = _free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: "ok", 1: "ok"}


def test_single_process_is_identity():
    from timegan_b200 import dist as D
    assert not D.is_enabled() and D.world_size() == 1 and D.rank() == 0
    t = torch.arange(4.0)
    s, n = D.allreduce_stats(t, 7)
    assert s is t and n == 7.0
    assert torch.equal(D.shard_batch(t), t)
    assert D.global_mean(t.sum(), 4).item() == 1.5
