"""Generates tests/golden/vdb_bonsai_small.json from the reference's own data file
(/root/reference/data/vdb/bonsai_small.vdb, the one real volume payload in the checkout)
through libcvr_b200.so's reader.  Run in the build container, where the reference is mounted:
    python tests/golden/make_vdb_golden.py
The INDEPENDENT pins are the file's own metadata written by OpenVDB (file_voxel_count,
file_bbox_min/max, per-grid stream offsets, which the reader checks while parsing) and the
converter's arithmetic (scripts/convert-mhd/mhd_to_vdb.py:47-63: density = smoothstep of the
min-max normalised scan, albedo = (density, 0, 0)); the digests pin the decoded values."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cudavolumerenderer_b200.vdb import VdbFile  # noqa: E402

SRC = "/root/reference/data/vdb/bonsai_small.vdb"


def main():
    out = {"source": SRC, "file_sha256": hashlib.sha256(open(SRC, "rb").read()).hexdigest(), "grids": {}}
    with VdbFile(SRC) as f:
        for g in f.grids():
            name = g["name"]
            dense = f.densify(name)
            org, msk, val = f.leaves(name)
            out["grids"][name] = {
                "info": {k: (list(v) if isinstance(v, tuple) else v) for k, v in g.items()},
                "file_voxel_count": int(f.meta(name, "file_voxel_count")),
                "file_bbox_min": [int(v) for v in f.meta(name, "file_bbox_min").split()],
                "file_bbox_max": [int(v) for v in f.meta(name, "file_bbox_max").split()],
                "file_compression": f.meta(name, "file_compression"),
                "dense_sha256": hashlib.sha256(np.ascontiguousarray(dense).tobytes()).hexdigest(),
                "dense_sum": float(dense.astype(np.float64).sum()),
                "dense_max": float(dense.max()),
                "leaf_origin_sha256": hashlib.sha256(org.tobytes()).hexdigest(),
                "leaf_mask_sha256": hashlib.sha256(msk.tobytes()).hexdigest(),
                "leaf_value_sha256": hashlib.sha256(np.ascontiguousarray(val).tobytes()).hexdigest(),
            }
    with open(os.path.join(ROOT, "tests", "golden", "vdb_bonsai_small.json"), "w") as fp:
        json.dump(out, fp, indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True)[:1500])


if __name__ == "__main__":
    main()
