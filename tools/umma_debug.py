"""Sweep the tcgen05 conv kernel over ResNet shapes and print an error table (GPU box diagnostic)."""
import json
import os
import sys
import traceback
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import numpy as np

from defer_b200 import _cabi as A

lib = A.load()
import torch  # noqa: E402

import test_gpu_kernels as T  # noqa: E402

rows = []
shapes = T.RESNET_SHAPES if len(sys.argv) < 2 else T.RESNET_SHAPES[: int(sys.argv[1])]
for fmt in ["bf16", "bf16x2"]:
    for i, sh in enumerate(shapes):
        n, h, w, cin, cout, k, s, pad = sh
        try:
            err, y, ref = T._conv_case(torch, lib, fmt, 2, n, h, w, cin, cout, k, s, pad, relu=(i % 2 == 0), residual=(i % 3 == 0), seed=i)
            d = np.abs(y - ref)
            bad = d > 1e-2 * np.abs(ref).max()
            info = {"fmt": fmt, "shape": sh, "err": err, "bad_frac": float(bad.mean())}
            if bad.any():
                idx = np.argwhere(bad)
                info["bad_first"] = idx[0].tolist()
                info["bad_rows"] = int(np.unique(idx[:, 1] * y.shape[2] + idx[:, 2]).size)
                info["bad_cols"] = int(np.unique(idx[:, 3]).size)
                info["nan"] = int(np.isnan(y).sum())
            print(json.dumps(info), flush=True)
            rows.append(info)
        except Exception as e:
            traceback.print_exc()
            print(json.dumps({"fmt": fmt, "shape": sh, "exception": str(e)}), flush=True)
            rows.append({"fmt": fmt, "shape": sh, "exception": str(e)})
            if "CUDA" in str(e) or "cuda" in str(e):
                break
os.makedirs(ROOT / "gpurun_out", exist_ok=True)
json.dump(rows, open(ROOT / "gpurun_out" / "umma_debug.json", "w"), indent=1)
