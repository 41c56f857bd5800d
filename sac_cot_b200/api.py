"""Host-side interface of the SAC-COT hot path over the C ABI (include/sac_cot.h).

The reference exposes no operator/plugin interface to mirror (/root/reference/README.md:1-2
is the entire repository), so this module mirrors the boundary BASELINE.json `north_star`
names — `sac_cot_register(src, dst, N, params, &R, &t, &inliers)` — with the same names,
argument meaning and error behaviour as the header.

`Registrar` wraps one `sac_cot_ctx` of a library that implements the ABI.  The module-level
helpers (`register`, `register_batch`, ...) always use the CUDA product library and raise if
it is missing or no GPU is usable: there is no CPU fallback in the product.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import _abi

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libsaccot.so")
_lib = None


class SacCotError(RuntimeError):
    def __init__(self, status: int, where: str, lib=None):
        msg = lib.sac_cot_strerror(status).decode() if lib is not None else ""
        super().__init__(f"{where} failed: status {status} ({msg})")
        self.status = status


def load_library(path: str | None = None) -> C.CDLL:
    """Load and bind the CUDA product library.  Fails loudly if it has not been built."""
    global _lib
    if path is None:
        if _lib is not None:
            return _lib
        if not os.path.exists(_LIB_PATH):
            raise ImportError(
                f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(sac_cot_b200 has no CPU fallback)")
        _lib = _abi.bind(C.CDLL(_LIB_PATH))
        return _lib
    return _abi.bind(C.CDLL(path))


@dataclass
class Result:
    R: np.ndarray        # (B,3,3) float32
    t: np.ndarray        # (B,3)   float32
    inliers: np.ndarray  # (B,)    int32


def _pack(pairs_src, pairs_dst):
    ns = [int(np.asarray(s).shape[0]) for s in pairs_src]
    offsets = np.zeros(len(ns) + 1, dtype=np.int64)
    np.cumsum(ns, out=offsets[1:])
    nd = [int(np.asarray(d).shape[0]) for d in pairs_dst]
    if ns != nd:
        raise ValueError("src and dst must hold the same number of points per pair")
    src = np.ascontiguousarray(np.concatenate([np.asarray(s, dtype=np.float32).reshape(-1, 3) for s in pairs_src]))
    dst = np.ascontiguousarray(np.concatenate([np.asarray(d, dtype=np.float32).reshape(-1, 3) for d in pairs_dst]))
    return src, dst, offsets


class Registrar:
    """One sac_cot_ctx.  `lib` defaults to the CUDA product library."""

    def __init__(self, lib: C.CDLL | None = None, device: int = 0, stream: int | None = None, **params):
        self.lib = lib if lib is not None else load_library()
        self._ctx = C.c_void_p()
        # stream: None -> private stream; 0 (torch's legacy default stream) -> cudaStreamLegacy handle 0x1
        handle = 0 if stream is None else (1 if int(stream) == 0 else int(stream))
        rc = self.lib.sac_cot_ctx_create(C.byref(self._ctx), int(device), C.c_void_p(handle))
        if rc != _abi.OK:
            self._ctx = C.c_void_p()
            raise SacCotError(rc, "sac_cot_ctx_create", self.lib)
        self.params = _abi.default_params(self.lib, **params)

    # -- lifetime -------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self.lib.sac_cot_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- knobs ----------------------------------------------------------------------
    def set(self, name: str, value: int):
        rc = self.lib.sac_cot_ctx_set(self._ctx, name.encode(), int(value))
        if rc != _abi.OK:
            raise SacCotError(rc, f"sac_cot_ctx_set({name})", self.lib)

    def get(self, name: str) -> int:
        v = C.c_int64()
        rc = self.lib.sac_cot_ctx_get(self._ctx, name.encode(), C.byref(v))
        if rc != _abi.OK:
            raise SacCotError(rc, f"sac_cot_ctx_get({name})", self.lib)
        return int(v.value)

    # -- registration ---------------------------------------------------------------
    def register(self, src, dst) -> tuple[np.ndarray, np.ndarray, int]:
        res = self.register_batch([src], [dst])
        return res.R[0], res.t[0], int(res.inliers[0])

    def register_batch(self, pairs_src, pairs_dst) -> Result:
        """Independent pairs given as sequences of (N_b,3) arrays (host)."""
        B = len(pairs_src)
        if B != len(pairs_dst):
            raise ValueError("need as many dst as src arrays")
        if B == 0:
            return Result(np.zeros((0, 3, 3), np.float32), np.zeros((0, 3), np.float32), np.zeros(0, np.int32))
        src, dst, offsets = _pack(pairs_src, pairs_dst)
        return self.register_packed(src, dst, offsets)

    def register_packed(self, src: np.ndarray, dst: np.ndarray, offsets: np.ndarray) -> Result:
        """Pairs packed back to back in host arrays; see sac_cot_register_packed."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        src = np.ascontiguousarray(src, dtype=np.float32).reshape(-1, 3)
        dst = np.ascontiguousarray(dst, dtype=np.float32).reshape(-1, 3)
        if offsets.ndim != 1 or len(offsets) < 1 or offsets[0] != 0 or not (len(src) == len(dst) == offsets[-1]):
            raise ValueError("offsets must start at 0 and end at the number of points of src and of dst")
        B = len(offsets) - 1
        R = np.empty((B, 3, 3), np.float32)
        t = np.empty((B, 3), np.float32)
        inl = np.empty(B, np.int32)
        rc = self.lib.sac_cot_register_packed(
            self._ctx, src.ctypes.data, dst.ctypes.data, offsets.ctypes.data_as(C.POINTER(C.c_int64)), B,
            C.byref(self.params), R.ctypes.data, t.ctypes.data, inl.ctypes.data, _abi.LOC_HOST)
        if rc != _abi.OK:
            raise SacCotError(rc, "sac_cot_register_packed", self.lib)
        return Result(R, t, inl)

    def register_packed_ptr(self, src_ptr: int, dst_ptr: int, offsets: np.ndarray, R_ptr: int, t_ptr: int,
                            inl_ptr: int, location: int):
        """Raw-pointer form (host or device buffers); enqueue-only for LOC_DEVICE."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        rc = self.lib.sac_cot_register_packed(
            self._ctx, src_ptr, dst_ptr, offsets.ctypes.data_as(C.POINTER(C.c_int64)), len(offsets) - 1,
            C.byref(self.params), R_ptr, t_ptr, inl_ptr, location)
        if rc != _abi.OK:
            raise SacCotError(rc, "sac_cot_register_packed", self.lib)

    def register_pointer_batch(self, pairs_src, pairs_dst) -> Result:
        """Same as register_batch but through sac_cot_register_batch (pointer arrays)."""
        B = len(pairs_src)
        srcs = [np.ascontiguousarray(s, dtype=np.float32).reshape(-1, 3) for s in pairs_src]
        dsts = [np.ascontiguousarray(d, dtype=np.float32).reshape(-1, 3) for d in pairs_dst]
        Ns = np.array([s.shape[0] for s in srcs], dtype=np.int32)
        sp = (C.POINTER(C.c_float) * B)(*[_abi.fptr(s) for s in srcs])
        dp = (C.POINTER(C.c_float) * B)(*[_abi.fptr(d) for d in dsts])
        R = np.empty((B, 3, 3), np.float32)
        t = np.empty((B, 3), np.float32)
        inl = np.empty(B, np.int32)
        rc = self.lib.sac_cot_register_batch(self._ctx, sp, dp, Ns.ctypes.data_as(C.POINTER(C.c_int32)), B,
                                             C.byref(self.params), _abi.fptr(R), _abi.fptr(t),
                                             inl.ctypes.data_as(C.POINTER(C.c_int32)))
        if rc != _abi.OK:
            raise SacCotError(rc, "sac_cot_register_batch", self.lib)
        return Result(R, t, inl)

    # -- correspondence front end --------------------------------------
    def match_packed_ptr(self, desc_src: int, xyz_src: int, offs_src: np.ndarray, desc_dst: int, xyz_dst: int,
                         offs_dst: np.ndarray, dim: int, nn: int, corr_src: int, corr_dst: int, location: int):
        """Raw-pointer form of sac_cot_match_packed (host or device buffers); enqueue-only for LOC_DEVICE."""
        offs_src = np.ascontiguousarray(offs_src, dtype=np.int64)
        offs_dst = np.ascontiguousarray(offs_dst, dtype=np.int64)
        if len(offs_src) != len(offs_dst):
            raise ValueError("offs_src and offs_dst must describe the same number of pairs")
        rc = self.lib.sac_cot_match_packed(
            self._ctx, desc_src, xyz_src, offs_src.ctypes.data_as(C.POINTER(C.c_int64)), desc_dst, xyz_dst,
            offs_dst.ctypes.data_as(C.POINTER(C.c_int64)), len(offs_src) - 1, int(dim), nn, corr_src, corr_dst, location)
        if rc != _abi.OK:
            raise SacCotError(rc, "sac_cot_match_packed", self.lib)

    def match_batch(self, desc_src, xyz_src, desc_dst, xyz_dst):
        """Nearest-neighbour matching of descriptors, one entry per pair: sequences of (Ns_b, dim) / (Ns_b, 3) source and
        (Nd_b, dim) / (Nd_b, 3) target arrays (host).  Returns (nn, corr_src, corr_dst, offsets): nn[i] is the matched
        target keypoint (index local to the pair) of source keypoint i; corr_src / corr_dst / offsets feed
        register_packed as they are."""
        B = len(desc_src)
        if not (B == len(xyz_src) == len(desc_dst) == len(xyz_dst)) or B == 0:
            raise ValueError("need the same (non-zero) number of pairs in every argument")
        ds = [np.ascontiguousarray(a, dtype=np.float32) for a in desc_src]
        dd = [np.ascontiguousarray(a, dtype=np.float32) for a in desc_dst]
        dim = ds[0].shape[1]
        if any(a.ndim != 2 or a.shape[1] != dim for a in ds + dd):
            raise ValueError("descriptors must be (rows, dim) arrays of one width")
        xs = [np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 3) for a in xyz_src]
        xd = [np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 3) for a in xyz_dst]
        if any(len(a) != len(b) for a, b in zip(ds, xs)) or any(len(a) != len(b) for a, b in zip(dd, xd)):
            raise ValueError("every keypoint needs a descriptor")
        os_ = np.zeros(B + 1, np.int64)
        od_ = np.zeros(B + 1, np.int64)
        np.cumsum([len(a) for a in ds], out=os_[1:])
        np.cumsum([len(a) for a in dd], out=od_[1:])
        DS, DD, XS, XD = (np.ascontiguousarray(np.concatenate(v)) for v in (ds, dd, xs, xd))
        nn = np.empty(int(os_[-1]), np.int32)
        cs = np.empty((int(os_[-1]), 3), np.float32)
        cd = np.empty((int(os_[-1]), 3), np.float32)
        self.match_packed_ptr(DS.ctypes.data, XS.ctypes.data, os_, DD.ctypes.data, XD.ctypes.data, od_, dim, nn.ctypes.data,
                              cs.ctypes.data, cd.ctypes.data, _abi.LOC_HOST)
        return nn, cs, cd, os_

    def match_mutual_batch(self, desc_src, xyz_src, desc_dst, xyz_dst):
        """Matching in both directions + the mutual-nearest-neighbour filter (sac_cot_match_mutual): returns
        (corr_src, corr_dst, offsets) holding, per pair and in order, the correspondences i with
        nn_back[nn[i]] == i.  Pairs left with fewer than three correspondences cannot be registered."""
        nn, cs, cd, os_ = self.match_batch(desc_src, xyz_src, desc_dst, xyz_dst)
        nb, _, _, od_ = self.match_batch(desc_dst, xyz_dst, desc_src, xyz_src)
        out_s = np.empty_like(cs)
        out_d = np.empty_like(cd)
        out_off = np.zeros(len(os_), np.int64)
        rc = self.lib.sac_cot_match_mutual(
            self._ctx, nn.ctypes.data, nb.ctypes.data, cs.ctypes.data, cd.ctypes.data,
            os_.ctypes.data_as(C.POINTER(C.c_int64)), od_.ctypes.data_as(C.POINTER(C.c_int64)), len(os_) - 1,
            out_s.ctypes.data, out_d.ctypes.data, out_off.ctypes.data, _abi.LOC_HOST)
        if rc != _abi.OK:
            raise SacCotError(rc, "sac_cot_match_mutual", self.lib)
        n = int(out_off[-1])
        return out_s[:n], out_d[:n], out_off

    def match(self, desc_src, xyz_src, desc_dst, xyz_dst):
        """One pair: (nn, corr_src, corr_dst)."""
        nn, cs, cd, _ = self.match_batch([desc_src], [xyz_src], [desc_dst], [xyz_dst])
        return nn, cs, cd

    # -- sharded single pair ------------------------------------------
    def sharded_phase1(self, src, dst, rank: int, world: int):
        src = np.ascontiguousarray(src, dtype=np.float32).reshape(-1, 3)
        dst = np.ascontiguousarray(dst, dtype=np.float32).reshape(-1, 3)
        N = src.shape[0]
        t_partial = np.zeros(N, np.uint64)
        cand = np.zeros(self.params.num_edges, np.uint64)
        rc = self.lib.sac_cot_sharded_phase1(self._ctx, _abi.fptr(src), _abi.fptr(dst), N, C.byref(self.params),
                                             rank, world, t_partial.ctypes.data_as(C.POINTER(C.c_uint64)),
                                             cand.ctypes.data_as(C.POINTER(C.c_uint64)))
        if rc != _abi.OK:
            raise SacCotError(rc, "sac_cot_sharded_phase1", self.lib)
        return t_partial, cand

    def sharded_phase2(self, t_all: np.ndarray, cand_all: np.ndarray) -> int:
        t_all = np.ascontiguousarray(t_all, dtype=np.uint64)
        cand_all = np.ascontiguousarray(cand_all, dtype=np.uint64)
        best = C.c_uint64()
        rc = self.lib.sac_cot_sharded_phase2(self._ctx, t_all.ctypes.data_as(C.POINTER(C.c_uint64)),
                                             cand_all.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(best))
        if rc != _abi.OK:
            raise SacCotError(rc, "sac_cot_sharded_phase2", self.lib)
        return int(best.value)

    def sharded_phase3(self, best_key: int):
        R = np.empty((3, 3), np.float32)
        t = np.empty(3, np.float32)
        inl = C.c_int32()
        rc = self.lib.sac_cot_sharded_phase3(self._ctx, C.c_uint64(best_key), _abi.fptr(R), _abi.fptr(t), C.byref(inl))
        if rc != _abi.OK:
            raise SacCotError(rc, "sac_cot_sharded_phase3", self.lib)
        return R, t, int(inl.value)

    # -- in-library collectives (GPU library: NCCL communicator held by the ctx) -------------
    def comm_init(self, group=None, rank: int | None = None, world: int | None = None, unique_id: bytes | None = None):
        """Creates the ctx's NCCL communicator (sac_cot_ctx_comm_init; collective over the ranks).

        With a torch.distributed process group (default: the world group) rank 0 draws the unique id
        (sac_cot_comm_unique_id) and broadcasts its 128 bytes over that group — any backend will do, the id is
        plain data.  Without torch.distributed pass rank, world and the id obtained on rank 0."""
        if unique_id is None:
            import torch
            import torch.distributed as dist

            rank, world = dist.get_rank(group), dist.get_world_size(group)
            buf = (C.c_ubyte * _abi.COMM_ID_BYTES)()
            if rank == 0:
                rc = self.lib.sac_cot_comm_unique_id(buf)
                if rc != _abi.OK:
                    raise SacCotError(rc, "sac_cot_comm_unique_id", self.lib)
            dev = torch.device("cuda", self.get("device")) if dist.get_backend(group) == "nccl" else torch.device("cpu")
            tid = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=dev)
            dist.broadcast(tid, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            unique_id = bytes(tid.cpu().tolist())
        if len(unique_id) != _abi.COMM_ID_BYTES:
            raise ValueError(f"unique_id must be {_abi.COMM_ID_BYTES} bytes")
        rc = self.lib.sac_cot_ctx_comm_init(self._ctx, unique_id, int(rank), int(world))
        if rc != _abi.OK:
            raise SacCotError(rc, "sac_cot_ctx_comm_init", self.lib)

    def set_comm(self, nccl_comm: int, rank: int, world: int):
        """Adopts an existing ncclComm_t (raw handle); 0 detaches."""
        rc = self.lib.sac_cot_ctx_set_comm(self._ctx, C.c_void_p(nccl_comm), rank, world)
        if rc != _abi.OK:
            raise SacCotError(rc, "sac_cot_ctx_set_comm", self.lib)

    def register_sharded_ptr(self, src_ptr: int, dst_ptr: int, N: int, R_ptr: int, t_ptr: int, inl_ptr: int,
                             location: int):
        """Raw-pointer form of sac_cot_register_sharded (host or device buffers); enqueue-only for LOC_DEVICE."""
        rc = self.lib.sac_cot_register_sharded(self._ctx, src_ptr, dst_ptr, int(N), C.byref(self.params), R_ptr, t_ptr,
                                               inl_ptr, location)
        if rc != _abi.OK:
            raise SacCotError(rc, "sac_cot_register_sharded", self.lib)

    def register_sharded(self, src, dst, group=None):
        """One large pair split over the ranks: exactly two exchanges — all-gather of the
        per-node partial triangle sums and each rank's top-K_e edge candidates, then all-reduce(max) of the
        packed (score, hypothesis id) key.  Every rank returns the same (R, t, inliers).

        If the ctx holds a communicator (comm_init / set_comm) this is a call-through to
        sac_cot_register_sharded: the library enqueues kernels and NCCL collectives on its own stream, device to
        device.  Otherwise the three phase calls are driven from here and torch.distributed carries the two
        exchanges (gloo in the CPU tests, where the compute runs on the oracle)."""
        src = np.ascontiguousarray(src, dtype=np.float32).reshape(-1, 3)
        dst = np.ascontiguousarray(dst, dtype=np.float32).reshape(-1, 3)
        if src.shape != dst.shape:
            raise ValueError("src and dst must have the same shape")
        if self.get("device") >= 0 and self.get("comm_world") > 0:
            R = np.empty((3, 3), np.float32)
            t = np.empty(3, np.float32)
            inl = np.zeros(1, np.int32)
            self.register_sharded_ptr(src.ctypes.data, dst.ctypes.data, src.shape[0], R.ctypes.data, t.ctypes.data,
                                      inl.ctypes.data, _abi.LOC_HOST)
            return R, t, int(inl[0])
        import torch
        import torch.distributed as dist

        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        t_partial, cand = self.sharded_phase1(src, dst, rank, world)
        dev = torch.device("cuda", self.get("device")) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        # u64 payloads travel as int64 bit patterns (same width; all-gather is a pure copy)
        payload = torch.from_numpy(np.concatenate([t_partial, cand]).view(np.int64)).to(dev)
        gathered = torch.empty(world * payload.numel(), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(gathered, payload, group=group)                      # exchange #1
        g = gathered.cpu().numpy().view(np.uint64).reshape(world, -1)
        N = t_partial.shape[0]
        best = self.sharded_phase2(g[:, :N], g[:, N:])
        # keys are < 2^63 (score < 2^47), so signed max == unsigned max
        key = torch.tensor([best], dtype=torch.int64, device=dev)
        dist.all_reduce(key, op=dist.ReduceOp.MAX, group=group)                          # exchange #2
        return self.sharded_phase3(int(key.item()))

    # -- parity access ----------------------------------------------------------------
    def debug(self, pair: int, which: int) -> np.ndarray:
        need = C.c_size_t()
        rc = self.lib.sac_cot_debug_get(self._ctx, pair, which, None, 0, C.byref(need))
        if rc not in (_abi.OK, _abi.E_CAPACITY):
            raise SacCotError(rc, f"sac_cot_debug_get({which})", self.lib)
        dt = np.dtype(_abi.DBG_DTYPES[which])
        out = np.empty(need.value // dt.itemsize, dtype=dt)
        if need.value:
            rc = self.lib.sac_cot_debug_get(self._ctx, pair, which, out.ctypes.data, out.nbytes, C.byref(need))
            if rc != _abi.OK:
                raise SacCotError(rc, f"sac_cot_debug_get({which})", self.lib)
        return out


class Group:
    """sac_cot_group: one batch over several GPUs of the box, pair b on device b mod G, no communication."""

    def __init__(self, devices, lib: C.CDLL | None = None, **params):
        self.lib = lib if lib is not None else load_library()
        self._grp = C.c_void_p()
        devs = np.ascontiguousarray(list(devices), dtype=np.int32)
        rc = self.lib.sac_cot_group_create(C.byref(self._grp), devs.ctypes.data_as(C.POINTER(C.c_int32)), len(devs))
        if rc != _abi.OK:
            self._grp = C.c_void_p()
            raise SacCotError(rc, "sac_cot_group_create", self.lib)
        self.params = _abi.default_params(self.lib, **params)

    def close(self):
        if getattr(self, "_grp", None) is not None and self._grp:
            self.lib.sac_cot_group_destroy(self._grp)
            self._grp = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __len__(self):
        return int(self.lib.sac_cot_group_size(self._grp))

    def set(self, name: str, value: int):
        rc = self.lib.sac_cot_group_set(self._grp, name.encode(), int(value))
        if rc != _abi.OK:
            raise SacCotError(rc, f"sac_cot_group_set({name})", self.lib)

    def get(self, index: int, name: str) -> int:
        """sac_cot_ctx_get on member `index`."""
        ctx = self.lib.sac_cot_group_ctx(self._grp, index)
        v = C.c_int64()
        rc = self.lib.sac_cot_ctx_get(ctx, name.encode(), C.byref(v))
        if rc != _abi.OK:
            raise SacCotError(rc, f"sac_cot_ctx_get({name})", self.lib)
        return int(v.value)

    def register_packed_ptr(self, src_ptr: int, dst_ptr: int, offsets: np.ndarray, R_ptr: int, t_ptr: int, inl_ptr: int):
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        rc = self.lib.sac_cot_group_register_packed(
            self._grp, src_ptr, dst_ptr, offsets.ctypes.data_as(C.POINTER(C.c_int64)), len(offsets) - 1,
            C.byref(self.params), R_ptr, t_ptr, inl_ptr)
        if rc != _abi.OK:
            raise SacCotError(rc, "sac_cot_group_register_packed", self.lib)

    def register_batch(self, pairs_src, pairs_dst) -> Result:
        B = len(pairs_src)
        if B != len(pairs_dst):
            raise ValueError("need as many dst as src arrays")
        if B == 0:
            return Result(np.zeros((0, 3, 3), np.float32), np.zeros((0, 3), np.float32), np.zeros(0, np.int32))
        src, dst, offsets = _pack(pairs_src, pairs_dst)
        R = np.empty((B, 3, 3), np.float32)
        t = np.empty((B, 3), np.float32)
        inl = np.empty(B, np.int32)
        self.register_packed_ptr(src.ctypes.data, dst.ctypes.data, offsets, R.ctypes.data, t.ctypes.data, inl.ctypes.data)
        return Result(R, t, inl)


# ---- module-level convenience on the CUDA product library ---------------------------------
def register(src, dst, **params):
    """sac_cot_register: one pair on the process-global ctx of the CUDA library (device 0)."""
    lib = load_library()
    p = _abi.default_params(lib, **params)
    src = np.ascontiguousarray(src, dtype=np.float32).reshape(-1, 3)
    dst = np.ascontiguousarray(dst, dtype=np.float32).reshape(-1, 3)
    if src.shape != dst.shape:
        raise ValueError("src and dst must have the same shape")
    R = np.empty((3, 3), np.float32)
    t = np.empty(3, np.float32)
    inl = C.c_int32()
    rc = lib.sac_cot_register(_abi.fptr(src), _abi.fptr(dst), src.shape[0], C.byref(p), _abi.fptr(R), _abi.fptr(t),
                              C.byref(inl))
    if rc != _abi.OK:
        raise SacCotError(rc, "sac_cot_register", lib)
    return R, t, int(inl.value)


def register_batch(pairs_src, pairs_dst, device: int = 0, **params) -> Result:
    with Registrar(device=device, **params) as reg:
        return reg.register_batch(pairs_src, pairs_dst)
