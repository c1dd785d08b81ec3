"""include/lis.h is a plain C header: it compiles as C99, and a C program can link the library and use
the host-only entry points without any Python or C++ in between."""
import os
import shutil
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent

C_PROG = r"""
#include <stdio.h>
#include <string.h>
#include "lis.h"

int main(void) {
  int32_t lens[3] = {20, 150, 0};
  int32_t seg_query[8], seg_lo[8], seg_hi[8], mt_seg[8];
  int64_t n_mt = 0;
  int64_t n = lis_plan_queries(lens, 3, 8, seg_query, seg_lo, seg_hi, 8, mt_seg, &n_mt);
  /* 20 | 150 cut at rows 64 and 128: [0,20) [20,64) [64,128) [128,170) */
  if (n != 4 || n_mt != 2) { printf("bad plan %lld %lld\n", (long long)n, (long long)n_mt); return 1; }
  if (seg_lo[1] != 20 || seg_hi[1] != 64 || seg_lo[2] != 64 || seg_hi[2] != 128 || seg_lo[3] != 128 || seg_hi[3] != 170) return 2;
  if (lis_abi_version() != LIS_ABI_VERSION) return 3;
  if (lis_set_tuning(100, 0, 0, 0, 0) != LIS_E_INVALID) return 4;
  if (strstr(lis_last_error(), "tile_n") == NULL) return 5;
  if (lis_topk_workspace_bytes(1, 100, 10) != 256) return 6;
  printf("c-abi ok\n");
  return 0;
}
"""


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_header_is_c99_and_library_links_from_c(tmp_path):
    import importlib

    native = importlib.import_module("multi-modal_colpali_b200._native")
    native.load()
    lib = native.lib_path()
    src = tmp_path / "abi.c"
    src.write_text(C_PROG)
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", f"-I{ROOT / 'include'}", str(src), "-o", str(exe),
                    str(lib), f"-Wl,-rpath,{lib.parent}"], check=True, capture_output=True, text=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "c-abi ok" in r.stdout, (r.returncode, r.stdout, r.stderr)
