#!/usr/bin/env python3
"""Pack the reference's DATA files (not sources) into this repo's asset layout.

Run once in the build container (where /root/reference is mounted); the outputs are
committed because /root/reference does not exist on the GPU box.

  * NN weights: cpp/NNmodel/{self,env}/parameter/{weight,bias}_k.txt are 8-significant-digit
    decimal text.  They are parsed straight to IEEE double (Python float() is correctly
    rounded, like the strtod the reference's `ifstream >> double` uses,
    SelfCollisionModel.cpp:19-57) and stored as little-endian fp64 in one file per network:
        8 bytes  magic  b"MPCCNN1\\0"
        int32    n_layers
        int32    (out, in) per layer
        float64  W_l (out*in, row per output neuron) then b_l (out), layer after layer
  * Params: cpp/Params/{model,cost,bounds,normalization,sqp,track,config}.json are re-emitted
    with json.dump (same keys, same values) -- the loaders accept the reference's files unchanged.
"""
import json
import struct
import sys
from pathlib import Path

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/cpp")
OUT = Path(__file__).resolve().parent.parent / "mpcc_manipulator_b200" / "assets"


def read_txt(p):
    return [float(t) for t in p.read_text().split()]


def pack_nn(src, dims, dst):
    blob = bytearray(b"MPCCNN1\0")
    blob += struct.pack("<i", len(dims))
    for o, i in dims:
        blob += struct.pack("<ii", o, i)
    for k, (o, i) in enumerate(dims):
        w = read_txt(src / f"weight_{k}.txt")
        b = read_txt(src / f"bias_{k}.txt")
        assert len(w) == o * i and len(b) == o, (src, k, len(w), len(b))
        blob += struct.pack(f"<{len(w)}d", *w)
        blob += struct.pack(f"<{len(b)}d", *b)
    dst.write_bytes(bytes(blob))
    print(dst, len(blob), "bytes")


def main():
    (OUT / "nn").mkdir(parents=True, exist_ok=True)
    (OUT / "params").mkdir(parents=True, exist_ok=True)
    # architectures: osqp_interface.cpp:35-43
    pack_nn(REF / "NNmodel/self/parameter", [(256, 21), (64, 256), (1, 64)], OUT / "nn" / "self_collision.f64")
    pack_nn(REF / "NNmodel/env/parameter", [(256, 30), (256, 256), (256, 256), (256, 256), (9, 256)], OUT / "nn" / "env_collision.f64")
    for name in ["model", "cost", "bounds", "normalization", "sqp", "track", "config"]:
        d = json.loads((REF / "Params" / f"{name}.json").read_text())
        (OUT / "params" / f"{name}.json").write_text(json.dumps(d, indent=1) + "\n")
        print(OUT / "params" / f"{name}.json")


if __name__ == "__main__":
    main()
