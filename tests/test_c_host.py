"""The drop-in boundary as a C host sees it: a C99 translation unit that includes include/sac_cot.h, links against the
product library with nothing but plain pointers and runs `sac_cot_register` (the north_star signature).  Without a
GPU the call must refuse with SAC_COT_E_NODEVICE (there is no CPU fallback); with one it must return a pose."""
import os
import shutil
import subprocess

import pytest

from sac_cot_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "sac_cot_b200", "lib")

SRC = r"""
#include <stdio.h>
#include <stdlib.h>
#include "sac_cot.h"

int main(void) {
  enum { N = 64 };
  float src[N][3], dst[N][3], R[9], t[3];
  int32_t inliers = -1;
  sac_cot_params p;
  unsigned s = 12345u;
  int i, c, rc;
  if (sac_cot_params_default(&p) != SAC_COT_OK) return 2;
  for (i = 0; i < N; ++i)
    for (c = 0; c < 3; ++c) {
      s = s * 1664525u + 1013904223u;
      src[i][c] = (float)(s >> 8) / 16777216.0f * 3.0f;
      dst[i][c] = src[i][c] + (c == 0 ? 0.5f : 0.0f);   /* a pure translation: every pair is compatible */
    }
  rc = sac_cot_register(&src[0][0], &dst[0][0], N, &p, R, t, &inliers);
  printf("%d %d %.3f %.3f %.3f %s\n", rc, (int)inliers, t[0], t[1], t[2], sac_cot_strerror(rc));
  return 0;
}
"""


def _run_c_host(tmp_path):
    if not os.path.exists(os.path.join(LIBDIR, "libsaccot.so")):
        pytest.skip("product library not built")
    c = tmp_path / "host.c"
    c.write_text(SRC)
    exe = tmp_path / "host"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    str(c), "-o", str(exe), "-L", LIBDIR, "-lsaccot", f"-Wl,-rpath,{LIBDIR}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True, timeout=300).stdout.split()
    return int(out[0]), int(out[1]), [float(v) for v in out[2:5]]


@pytest.mark.skipif(shutil.which("gcc") is None, reason="no C compiler")
def test_c99_host_links_and_is_refused_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: see the gpu-marked test")
    rc, inliers, _ = _run_c_host(tmp_path)
    assert rc == _abi.E_NODEVICE and inliers in (-1, 0)


@pytest.mark.gpu
@pytest.mark.skipif(shutil.which("gcc") is None, reason="no C compiler")
def test_c99_host_registers_a_pure_translation(tmp_path):
    rc, inliers, t = _run_c_host(tmp_path)
    assert rc == _abi.OK and inliers == 64
    assert abs(t[0] - 0.5) < 1e-4 and abs(t[1]) < 1e-4 and abs(t[2]) < 1e-4
