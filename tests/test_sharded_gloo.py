"""world_size-2 (and 3) run of the sharded single-pair host logic over torch.distributed with the
gloo backend on CPU: Registrar.register_sharded performs exactly two exchanges (all-gather of
partial node sums + edge candidates, all-reduce(max) of the packed key).  The compute phases run
on the oracle here (the CUDA library needs a GPU); the GPU tests exercise the same phases on CUDA."""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, N, seed, out_dir):
    import torch.distributed as dist

    from sac_cot_b200 import _abi, synth
    from sac_cot_b200.api import Registrar

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    calls = {"all_gather": 0, "all_reduce": 0}
    orig_ag, orig_ar = dist.all_gather_into_tensor, dist.all_reduce

    def count_ag(*a, **k):
        calls["all_gather"] += 1
        return orig_ag(*a, **k)

    def count_ar(*a, **k):
        calls["all_reduce"] += 1
        return orig_ar(*a, **k)

    dist.all_gather_into_tensor, dist.all_reduce = count_ag, count_ar
    lib = _abi.bind(ctypes.CDLL(os.path.join(ROOT, "oracle", "libsaccot_oracle.so")))
    p = synth.make_pair(N, 0.05, seed)
    with Registrar(lib=lib) as reg:
        R, t, inl = reg.register_sharded(p.src, p.dst)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), R=R, t=t, inl=inl, ag=calls["all_gather"], ar=calls["all_reduce"])
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_register_sharded_over_gloo(oracle_lib, tmp_path, world):
    from sac_cot_b200 import synth
    from sac_cot_b200.api import Registrar

    N, seed = 1200, 4242
    mp.spawn(_worker, args=(world, _free_port(), N, seed, str(tmp_path)), nprocs=world, join=True)
    p = synth.make_pair(N, 0.05, seed)
    with Registrar(lib=oracle_lib) as reg:
        R0, t0, i0 = reg.register(p.src, p.dst)
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npz")
        np.testing.assert_array_equal(got["R"], R0)   # every rank returns the unsharded result, bit for bit
        np.testing.assert_array_equal(got["t"], t0)
        assert int(got["inl"]) == i0
        assert int(got["ag"]) == 1 and int(got["ar"]) == 1   # exactly two exchanges
