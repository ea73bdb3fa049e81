"""Peer-memory all-reduce (csrc/peer_allreduce.cu): host-side checks everywhere, the kernel itself on >= 2 GPUs."""
import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def test_peer_site_sizes_and_argument_checks():
    from timegan_b200 import _lib
    lib = _lib.lib
    chunk = lib.tg_peer_chunk_floats()
    assert chunk == 4096
    sizes = (C.c_longlong * 3)(1, chunk, chunk + 1)           # 1 + 1 + 2 chunks
    fb = C.c_size_t()
    db = lib.tg_peer_site_bytes(3, sizes, 8, C.byref(fb))
    assert db == 2 * 4 * chunk * 4 and fb.value == 4 * 8 * 4
    regions = (C.c_void_p * 2)(256, 256)
    ptrs = (C.c_void_p * 1)(256)
    one = (C.c_longlong * 1)(16)
    # never launches: every argument error is caught on the host
    assert lib.tg_peer_allreduce(None, 0, 1, regions, 0, 0, 256, 256, 1, ptrs, one) < 0       # world < 2
    assert "rank/world" in _lib.last_error()
    assert lib.tg_peer_allreduce(None, 0, 9, regions, 0, 0, 256, 256, 1, ptrs, one) < 0       # > 8 ranks
    assert lib.tg_peer_allreduce(None, 0, 2, regions, 0, 0, None, 256, 1, ptrs, one) < 0      # no epoch word
    assert lib.tg_peer_allreduce(None, 0, 2, regions, 8, 0, 256, 256, 1, ptrs, one) < 0       # misaligned staging
    assert "misaligned" in _lib.last_error()
    assert lib.tg_peer_allreduce(None, 0, 2, regions, 0, 0, 256, 256, 49, ptrs, one) < 0      # too many tensors


def _torchrun2(tool, port):
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                           "--master-addr", "127.0.0.1", "--master-port", str(port), str(ROOT / "tools" / tool)],
                          capture_output=True, text=True, env=env, timeout=400)


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs of one box")
def test_peer_allreduce_matches_nccl_two_ranks():
    r = _torchrun2("check_peer_allreduce.py", 29533)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count(": OK") == 2


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs of one box")
def test_dp_equals_single_process_and_graph_equals_eager_two_ranks():
    """DP on two ranks == one process on the global batch, and the DP step replayed from the CUDA
    graph (peer all-reduce kernels inside the graph) == the DP step issued eagerly."""
    r = _torchrun2("check_dp_equivalence.py", 29534)
    assert r.returncode == 0 and r.stdout.count("-> OK") == 2, r.stdout[-2000:] + r.stderr[-2000:]
    r = _torchrun2("check_dp_graph.py", 29535)
    assert r.returncode == 0 and r.stdout.count("-> OK") == 2, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs of one box")
def test_cli_under_torchrun_two_ranks(tmp_path):
    """`torchrun --nproc-per-node 2 -m timegan_b200.train_timegan ...`: each rank takes
    cuda:LOCAL_RANK, the peer transport comes up (so the default CUDA-graph path is used), a ragged epoch tail
    (13 windows, batch 4 -> 4,4,4,1: the last batch has fewer sequences than ranks and is skipped on both) trains
    through, rank 0 writes the artefacts, and both ranks end with bit-identical weights."""
    import numpy as np
    data = tmp_path / "preprocessed"
    data.mkdir()
    np.savez(data / "posture1_no_exo.npz", X=np.random.default_rng(3).random((13, 32, 14), dtype=np.float32))
    env = dict(os.environ, PYTHONPATH=str(ROOT), TIMEGAN_B200_DUMP_WEIGHT_SUM=str(tmp_path))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29536", "-m", "timegan_b200.train_timegan",
                        "--data_dir", str(data), "--out_dir", str(tmp_path / "runs"), "--batch_size", "4",
                        "--ae_epochs", "1", "--sup_epochs", "1", "--gan_steps", "9", "--layers", "2", "--z_dim", "8",
                        "--hidden_dim", "8", "--acf_max_lag", "8"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "Using device: cuda:0" in r.stdout and "Using device: cuda:1" in r.stdout
    run = tmp_path / "runs" / "posture1_no_exo"
    assert (run / "ckpt_latest.pt").exists() and (run / "synthetic.npz").exists()
    sums = [(tmp_path / f"weight_sum_rank{k}.txt").read_text() for k in (0, 1)]
    assert sums[0] == sums[1], sums
