"""
Regenerates the committed fixtures under tests/golden/ from the read-only
reference mount.  Run in the build container only (``/root/reference`` does not
exist on the GPU box; tests read the committed files, never the mount):

    python tests/golden/make_golden.py

Fixtures
--------
contours_pv.json           the one numeric golden vector the reference holds:
                           the printed ``cal_contours(121)`` array of
                           notebooks/1.Keff_atmos.ipynb (cell 3 output, file lines
                           102-119): first/last three fp32 levels of six
                           isentropic levels, 9 significant digits each (enough to
                           identify every fp32 value uniquely).
barotropic_vorticity.npz   lat[256], lon[512], absolute_vorticity[256,512] (fp32)
                           copied byte-for-byte out of
                           Data/barotropic_vorticity.nc (contiguous, unfiltered
                           HDF5 datasets at offsets 885 / 1909 / 10101,
                           SURVEY.md §8c) -- the input of tests/test_LWA.py.
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def notebook_contours():
    nb = json.load(open(os.path.join(REF, "notebooks", "1.Keff_atmos.ipynb")))
    for cell in nb["cells"]:
        for out in cell.get("outputs", []):
            text = "".join(out.get("text", []))
            if text.startswith("<xarray.DataArray 'pv' (level: 15, contour: 121)>"):
                body = text[text.index("array("):text.index("dtype=float32")]
                rows = []
                for line in body.split("]"):
                    nums = re.findall(r"-?\d\.\d+e[-+]\d+", line)
                    if len(nums) == 6:
                        rows.append(nums)
                return rows
    raise RuntimeError("golden output not found")


def main():
    rows = notebook_contours()
    assert len(rows) == 6, rows
    json.dump({"source": "notebooks/1.Keff_atmos.ipynb:102-119",
               "levels_N": 121, "increase": True, "dtype": "float32",
               "columns": [0, 1, 2, 118, 119, 120],
               "level_K": [265, 275, 285, 600, 700, 850],
               "printed": rows},
              open(os.path.join(HERE, "contours_pv.json"), "w"), indent=1)
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from oracle.xcontour_oracle import read_barotropic_vorticity
    lat, lon, q = read_barotropic_vorticity(
        os.path.join(REF, "Data", "barotropic_vorticity.nc"))
    assert abs(lat[0] + 89.46313) < 1e-4 and abs(lon[1] - 0.703125) < 1e-6
    assert not np.isnan(q).any()
    np.savez_compressed(os.path.join(HERE, "barotropic_vorticity.npz"),
                        latitude=lat, longitude=lon, absolute_vorticity=q)
    print("wrote fixtures to", HERE)


if __name__ == "__main__":
    main()
