"""Regenerate steppingstone_b200/data/{walker3d,mike}.json from the reference's MJCF files.

Run in the build container only (reads /root/reference):  python tools/compile_models.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from steppingstone_b200 import model  # noqa: E402

REF = os.environ.get("SS_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "mocca_envs", "data", "custom")

if __name__ == "__main__":
    os.makedirs(model.DATA_DIR, exist_ok=True)
    for kind, name in ((0, "walker3d"), (1, "mike")):
        T = model.compile_mjcf(os.path.join(SRC, name + ".xml"), kind)
        model.tables_to_json(T, os.path.join(model.DATA_DIR, name + ".json"))
        print(name, "total mass %.4f kg" % T["total_mass"], "breaking %.5f" % T["contact_breaking"])
        for i, n in enumerate(T["body_names"]):
            print("  %2d %-18s parent %2d m %.4f com %s" % (i, n, T["parent"][i], T["mass"][i], T["com"][i].round(4)))
