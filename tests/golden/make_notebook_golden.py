"""Extracts the ONE numeric output of the real reference that ships with its repository:
`examples/Birkinshaw_Gull_stacking.ipynb` cell 3 (`catalog.data.head()` right after
`Catalog("WebSky_lite")`), i.e. the reference's own `Catalog.build_dataframe`
(paint_bucket.py:308-364, astropy Planck18_arXiv_v2 underneath) run by its author on five WebSky
halos.  Inputs (x, y, z, v_x, v_y, v_z, M_200c, redshift) and 12 derived columns are printed with
6-9 significant digits; pandas hides lon / theta / phi / D_a behind "...".

Run in the build container only (reads /root/reference): writes tests/golden/notebook_catalog_head.json.
"""
import json
import os
import re

SRC = "/root/reference/examples/Birkinshaw_Gull_stacking.ipynb"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "notebook_catalog_head.json")


def main():
    nb = json.load(open(SRC))
    cell = nb["cells"][3]
    assert "".join(cell["source"]).strip() == "catalog.data.head()"
    text = "".join(cell["outputs"][0]["data"]["text/plain"])
    cols, rows = [], {i: [] for i in range(5)}
    for block in re.split(r"\n\s*\n", text):
        lines = [ln for ln in block.splitlines() if ln.strip() and not ln.startswith("[")]
        if not lines:
            continue
        names = [n for n in lines[0].replace("\\", " ").split() if n != "..."]
        cols += names
        for ln in lines[1:]:
            tok = [t for t in ln.replace("\\", " ").split() if t != "..."]
            rows[int(tok[0])] += tok[1:]
    table = {c: [float(rows[i][j]) for i in range(5)] for j, c in enumerate(cols)}
    # decimals pandas printed, per column: the rounding half-width is the tolerance floor of the test
    text_tok = {c: [rows[i][j] for i in range(5)] for j, c in enumerate(cols)}
    json.dump({"source": "examples/Birkinshaw_Gull_stacking.ipynb cell 3 (reference's own output)",
               "columns": cols, "values": table, "printed": text_tok}, open(OUT, "w"), indent=1)
    print(cols)


if __name__ == "__main__":
    main()
