"""Generate tests/golden fixtures from the oracle (fp64 executor, stored as fp32).

The reference cannot run here (TensorFlow absent), so these vectors pin the ORACLE and the
synthetic weight recipe, not the reference: parity stays "unpinned" (see oracle/keras_ref.py header).
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from defer_b200 import applications  # noqa: E402
from oracle import keras_ref as R  # noqa: E402

out = ROOT / "tests" / "golden"
out.mkdir(parents=True, exist_ok=True)
meta = {"weight_seed": 1, "input_seed": 0, "stride": 8, "channels": 16,
        "layers": ["activation", "max_pooling2d", "add_2", "add_6", "add_12", "add_15", "avg_pool"],
        "generator": "tools/make_golden.py", "executor": "oracle.keras_ref float64"}
m = applications.ResNet50(seed=meta["weight_seed"])
x = applications.synthetic_input(1, seed=meta["input_seed"])
wm = R.WireModel(m.to_json(), m.get_weights())
vals = wm.predict(x, dtype=np.float64, return_all=True)
arrays = {"probs": vals["fc1000"].astype(np.float32),
          "logits": wm.predict(x, dtype=np.float64, final_activation=False).astype(np.float32)}
for name in meta["layers"]:
    v = vals[name]
    if v.ndim == 4:
        arrays[name] = v[0, ::meta["stride"], ::meta["stride"], :meta["channels"]].astype(np.float32)
    else:
        arrays[name] = v[0, :meta["channels"]].astype(np.float32)
np.savez_compressed(out / "resnet50_seed1_input0.npz", **arrays)
(out / "meta.json").write_text(json.dumps(meta, indent=1))
print({k: v.shape for k, v in arrays.items()})
