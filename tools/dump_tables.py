"""Write the plain-C tables a non-Python host needs for ss_create() as binary blobs:
    model.bin (ss_model), params.bin (ss_params), mt_env.bin / mt_aux.bin ([n,625] uint32 MT19937 states, gym seeding).
usage: python tools/dump_tables.py OUTDIR [walker3d|mike] [n_envs] [seed]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from steppingstone_b200 import model, seeding

out = sys.argv[1]
kind = model.ENV_IDS[sys.argv[2] if len(sys.argv) > 2 else "walker3d"]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
seed = int(sys.argv[4]) if len(sys.argv) > 4 else 8
os.makedirs(out, exist_ok=True)
M = model.make_model(kind); P = model.make_params(kind)
open(os.path.join(out, "model.bin"), "wb").write(bytes(M))
open(os.path.join(out, "params.bin"), "wb").write(bytes(P))
seeding.env_streams(seed, 0, n).tofile(os.path.join(out, "mt_env.bin"))
seeding.aux_streams(kind, seed, 0, n).tofile(os.path.join(out, "mt_aux.bin"))
print("wrote %s: ss_model %d B, ss_params %d B, %d env streams" % (out, ctypes.sizeof(M), ctypes.sizeof(P), n))
