"""Correspondence dumps and registration-recall evaluation.

3DMatch / 3DLoMatch / KITTI are not available offline and the reference ships neither data nor loaders
(/root/reference/README.md:1-2), so this module fixes the on-disk layout the pipeline reads, in the two forms the
public registration benchmarks are usually exported in, and the metrics they are scored with:

  * correspondence form  (one .npz per pair)        src (N,3) f32, dst (N,3) f32            matched keypoints
                                                      [T_gt (4,4) f64]  [labels (N,) bool]    ground truth, inlier flags
  * descriptor form      (one .npz per pair)        xyz0 (Ns,3), feat0 (Ns,D), xyz1 (Nd,3), feat1 (Nd,D)  [T_gt]
                                                      keypoints + local descriptors of the two fragments (the layout of
                                                      per-fragment FPFH / FCGF exports, two fragments per file); these
                                                      go through the matching front end (sac_cot_match_packed) first
  * gt.log               3DMatch-style text: "i j n" followed by the 4x4 matrix, one block per pair

`evaluate` registers every pair of a directory with a Registrar (the CUDA library, or — in the CPU tests — the oracle)
and reports registration recall under the usual criteria (3DMatch: RE < 15 deg and TE < 0.30 m; KITTI: RE < 5 deg and
TE < 0.60 m), mean RE / TE over the successful pairs, the inlier ratio of the putative correspondences, and seconds
per pair.  `python -m sac_cot_b200.dumps synth|eval ...` is the command-line front.
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import time
from dataclasses import dataclass

import numpy as np

from . import synth

CRITERIA = {"3dmatch": (15.0, 0.30), "3dlomatch": (15.0, 0.30), "kitti": (5.0, 0.60)}


@dataclass
class Dump:
    name: str
    kind: str                      # "correspondences" | "descriptors"
    src: np.ndarray | None = None  # (N,3) matched keypoints (correspondence form)
    dst: np.ndarray | None = None
    xyz0: np.ndarray | None = None  # descriptor form
    feat0: np.ndarray | None = None
    xyz1: np.ndarray | None = None
    feat1: np.ndarray | None = None
    T_gt: np.ndarray | None = None  # (4,4) float64, dst ~= R src + t
    labels: np.ndarray | None = None


def write_correspondences(path: str, src, dst, T_gt=None, labels=None) -> None:
    arrs = {"src": np.asarray(src, np.float32).reshape(-1, 3), "dst": np.asarray(dst, np.float32).reshape(-1, 3)}
    if arrs["src"].shape != arrs["dst"].shape:
        raise ValueError("src and dst must hold the same number of points")
    if T_gt is not None:
        arrs["T_gt"] = np.asarray(T_gt, np.float64).reshape(4, 4)
    if labels is not None:
        arrs["labels"] = np.asarray(labels, bool).reshape(-1)
    np.savez(path, **arrs)


def write_descriptors(path: str, xyz0, feat0, xyz1, feat1, T_gt=None) -> None:
    arrs = {"xyz0": np.asarray(xyz0, np.float32).reshape(-1, 3), "feat0": np.asarray(feat0, np.float32),
            "xyz1": np.asarray(xyz1, np.float32).reshape(-1, 3), "feat1": np.asarray(feat1, np.float32)}
    if len(arrs["xyz0"]) != len(arrs["feat0"]) or len(arrs["xyz1"]) != len(arrs["feat1"]):
        raise ValueError("every keypoint needs a descriptor")
    if arrs["feat0"].ndim != 2 or arrs["feat1"].ndim != 2 or arrs["feat0"].shape[1] != arrs["feat1"].shape[1]:
        raise ValueError("descriptors must be (rows, dim) arrays of one width")
    if T_gt is not None:
        arrs["T_gt"] = np.asarray(T_gt, np.float64).reshape(4, 4)
    np.savez(path, **arrs)


def read_dump(path: str) -> Dump:
    with np.load(path) as z:
        keys = set(z.files)
        name = os.path.splitext(os.path.basename(path))[0]
        T = np.asarray(z["T_gt"], np.float64).reshape(4, 4) if "T_gt" in keys else None
        if {"src", "dst"} <= keys:
            src = np.ascontiguousarray(z["src"], np.float32).reshape(-1, 3)
            dst = np.ascontiguousarray(z["dst"], np.float32).reshape(-1, 3)
            if src.shape != dst.shape:
                raise ValueError(f"{path}: src and dst differ in length")
            return Dump(name, "correspondences", src=src, dst=dst, T_gt=T,
                        labels=np.asarray(z["labels"], bool) if "labels" in keys else None)
        if {"xyz0", "feat0", "xyz1", "feat1"} <= keys:
            return Dump(name, "descriptors", xyz0=np.ascontiguousarray(z["xyz0"], np.float32).reshape(-1, 3),
                        feat0=np.ascontiguousarray(z["feat0"], np.float32),
                        xyz1=np.ascontiguousarray(z["xyz1"], np.float32).reshape(-1, 3),
                        feat1=np.ascontiguousarray(z["feat1"], np.float32), T_gt=T)
    raise ValueError(f"{path}: neither a correspondence dump (src, dst) nor a descriptor dump (xyz0, feat0, xyz1, feat1)")


def read_gt_log(path: str) -> dict[tuple[int, int], np.ndarray]:
    """3DMatch-style gt.log: blocks of "i j n" + four rows of the 4x4 transform."""
    out = {}
    with open(path) as f:
        lines = [ln.split() for ln in f if ln.strip()]
    k = 0
    while k < len(lines):
        if len(lines) - k < 5 or len(lines[k]) < 2:
            raise ValueError(f"{path}: truncated block at line {k + 1}")
        T = np.array([[float(x) for x in lines[k + 1 + r]] for r in range(4)], np.float64)
        if T.shape != (4, 4):
            raise ValueError(f"{path}: block at line {k + 1} is not a 4x4 matrix")
        out[(int(lines[k][0]), int(lines[k][1]))] = T
        k += 5
    return out


def write_gt_log(path: str, poses: dict[tuple[int, int], np.ndarray], n: int = 0) -> None:
    with open(path, "w") as f:
        for (i, j), T in sorted(poses.items()):
            f.write(f"{i}\t{j}\t{n}\n")
            for r in range(4):
                f.write("\t".join(f"{v:.10e}" for v in np.asarray(T, np.float64)[r]) + "\n")


def iter_dumps(directory: str):
    for path in sorted(glob.glob(os.path.join(directory, "*.npz"))):
        yield read_dump(path)


def pose_errors(R, t, T_gt) -> tuple[float, float]:
    """(rotation error [deg], translation error [units]) of (R, t) against the 4x4 ground truth."""
    ang, dt = synth.pose_error(R, t, T_gt[:3, :3], T_gt[:3, 3])
    return float(np.degrees(ang)), dt


def evaluate(registrar, dumps, criterion: str = "3dmatch", tau: float | None = None) -> dict:
    """Registers every dump and scores it.  `tau` sets tau_compat = tau_inlier (default: the registrar's values)."""
    if criterion not in CRITERIA:
        raise ValueError(f"criterion must be one of {sorted(CRITERIA)}")
    re_max, te_max = CRITERIA[criterion]
    if tau is not None:
        registrar.params.tau_compat = registrar.params.tau_inlier = float(tau)
    tau_in = float(registrar.params.tau_inlier)
    rows = []
    for d in dumps:
        t0 = time.perf_counter()
        if d.kind == "descriptors":
            _, src, dst = registrar.match(d.feat0, d.xyz0, d.feat1, d.xyz1)
        else:
            src, dst = d.src, d.dst
        R, t, inl = registrar.register(src, dst)
        dt = time.perf_counter() - t0
        row = {"name": d.name, "kind": d.kind, "N": int(len(src)), "inliers": int(inl), "seconds": dt}
        if d.T_gt is not None:
            re, te = pose_errors(R, t, d.T_gt)
            resid = np.linalg.norm(src.astype(np.float64) @ d.T_gt[:3, :3].T + d.T_gt[:3, 3] - dst, axis=1)
            row.update(re_deg=re, te=te, success=bool(re < re_max and te < te_max), inlier_ratio=float((resid < tau_in).mean()))
        rows.append(row)
    scored = [r for r in rows if "success" in r]
    good = [r for r in scored if r["success"]]
    return {
        "criterion": criterion, "re_max_deg": re_max, "te_max": te_max, "pairs": len(rows), "pairs_with_ground_truth": len(scored),
        "recall": len(good) / len(scored) if scored else None,
        "mean_re_deg_success": float(np.mean([r["re_deg"] for r in good])) if good else None,
        "mean_te_success": float(np.mean([r["te"] for r in good])) if good else None,
        "mean_inlier_ratio": float(np.mean([r["inlier_ratio"] for r in scored])) if scored else None,
        "seconds_per_pair": float(np.mean([r["seconds"] for r in rows])) if rows else None,
        "per_pair": rows,
    }


def write_synthetic(directory: str, config: str, pairs: int, descriptors: bool = False, dim: int = 33) -> list[str]:
    """Synthetic stand-ins for the benchmark exports: pairs of BASELINE.json config `config` in dump form."""
    os.makedirs(directory, exist_ok=True)
    paths, poses = [], {}
    for b in range(pairs):
        p = synth.make_config_pair(config, b)
        T = np.eye(4)
        T[:3, :3], T[:3, 3] = p.R_gt, p.t_gt
        path = os.path.join(directory, f"pair_{b:04d}.npz")
        if descriptors:
            f, g = synth.make_descriptors(p, dim, seed=b)
            write_descriptors(path, p.src, f, p.dst, g, T)
        else:
            labels = np.zeros(len(p.src), bool)
            labels[p.inlier_idx] = True
            write_correspondences(path, p.src, p.dst, T, labels)
        poses[(b, b + 1)] = T
        paths.append(path)
    write_gt_log(os.path.join(directory, "gt.log"), poses, n=pairs + 1)
    return paths


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m sac_cot_b200.dumps")
    sub = ap.add_subparsers(dest="cmd", required=True)
    s = sub.add_parser("synth", help="write synthetic dumps of a BASELINE.json config")
    s.add_argument("--out", required=True)
    s.add_argument("--config", default="cfg2_3dmatch_256x5000", choices=sorted(synth.CONFIGS))
    s.add_argument("--pairs", type=int, default=8)
    s.add_argument("--descriptors", action="store_true")
    e = sub.add_parser("eval", help="register every dump of a directory and report recall")
    e.add_argument("--dir", required=True)
    e.add_argument("--criterion", default="3dmatch", choices=sorted(CRITERIA))
    e.add_argument("--tau", type=float, default=None, help="tau_compat = tau_inlier (default 0.1)")
    e.add_argument("--device", type=int, default=0)
    e.add_argument("--second-order", type=int, default=-1, help="so_min_common (>= 0 selects the second-order graph)")
    args = ap.parse_args(argv)
    if args.cmd == "synth":
        paths = write_synthetic(args.out, args.config, args.pairs, args.descriptors)
        print(json.dumps({"written": len(paths), "dir": args.out}))
        return 0
    from . import _abi
    from .api import Registrar
    with Registrar(device=args.device) as reg:   # the CUDA library; raises without a B200 (no CPU fallback)
        if args.second_order >= 0:
            reg.params.compat_mode = _abi.COMPAT_SECOND_ORDER
            reg.params.so_min_common = args.second_order
        rep = evaluate(reg, iter_dumps(args.dir), args.criterion, args.tau)
    rep.pop("per_pair")
    print(json.dumps(rep))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
