"""Graph-load timing for SURVEY section 8(f3): writes the GFA of a synthetic species (W lines) and times the host C++ reader against
ptx_upload_graph_gfa through `pantax-gpu-profile --time-gfa`.   python tools/bench_gfa.py [nodes] [paths]"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import synth
    from common import dataset_graphs

    nodes = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    paths = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    ds = synth.Dataset(20261019, [nodes], [paths])
    nl, ps, names = dataset_graphs(ds)[0]
    with tempfile.TemporaryDirectory() as d:
        fn = os.path.join(d, "1.gfa")
        with open(fn, "wb") as f:
            f.write(b"H\tVN:Z:1.1\n")
            ids = np.arange(1, len(nl) + 1)
            f.write("".join("S\t%d\t%s\n" % (i, "A" * int(l)) for i, l in zip(ids, nl)).encode())
            for n, p in zip(names, ps):
                f.write(("W\t%s\t0\tchr1\t0\t%d\t" % (n, len(p))).encode())
                f.write(np.char.add(">", (np.asarray(p, dtype=np.int64) + 1).astype(str)).astype("S").tobytes().replace(b"\x00", b""))
                f.write(b"\n")
        out = subprocess.run([os.path.join(ROOT, "pantax_b200", "pantax-gpu-profile"), "--time-gfa", fn], capture_output=True, text=True)
        print(out.stdout.strip() or out.stderr.strip())


if __name__ == "__main__":
    main()
