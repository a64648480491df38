"""Writes tests/golden/graph_fixture.{bin,bin.lz4,bin.zst,json}: one small Graph (types.rs:51-55) in the three on-disk
forms the reference reads (zip.rs:236-262).

The bincode stream is written here by hand with struct (bincode 1.3 default options: little endian, fixed-width integers,
u64 lengths; Vec<i64>, then BTreeMap<String, Vec<usize>> as count + entries in ascending key order) - independently of the
reader in pantax_b200/host/pantax_gpu_profile.cpp.  The LZ4 *frame* (what lz4_flex::frame::FrameEncoder writes, zip.rs:192-205)
and the zstd frame (zstd::Encoder, zip.rs:206-219) are produced by the system liblz4 / libzstd through ctypes.
Run once; the outputs are committed."""
import ctypes as C
import json
import os
import struct

HERE = os.path.dirname(os.path.abspath(__file__))


def bincode_graph(nodes_len, paths):
    out = struct.pack("<Q", len(nodes_len)) + b"".join(struct.pack("<q", v) for v in nodes_len)
    out += struct.pack("<Q", len(paths))
    for name in sorted(paths, key=lambda s: s.encode()):
        k = name.encode()
        out += struct.pack("<Q", len(k)) + k + struct.pack("<Q", len(paths[name])) + b"".join(struct.pack("<Q", v) for v in paths[name])
    return out


def lz4_frame(data: bytes) -> bytes:
    L = C.CDLL("liblz4.so.1")
    L.LZ4F_compressFrameBound.restype = C.c_size_t
    L.LZ4F_compressFrameBound.argtypes = [C.c_size_t, C.c_void_p]
    L.LZ4F_compressFrame.restype = C.c_size_t
    L.LZ4F_compressFrame.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
    cap = L.LZ4F_compressFrameBound(len(data), None)
    buf = C.create_string_buffer(cap)
    n = L.LZ4F_compressFrame(buf, cap, data, len(data), None)
    assert n <= cap
    return buf.raw[:n]


def zstd_frame(data: bytes) -> bytes:
    L = C.CDLL("libzstd.so.1")
    L.ZSTD_compressBound.restype = C.c_size_t
    L.ZSTD_compressBound.argtypes = [C.c_size_t]
    L.ZSTD_compress.restype = C.c_size_t
    L.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
    cap = L.ZSTD_compressBound(len(data))
    buf = C.create_string_buffer(cap)
    n = L.ZSTD_compress(buf, cap, data, len(data), 3)
    assert n <= cap
    return buf.raw[:n]


def main():
    import random

    rnd = random.Random(20261017)
    n = 3000
    nodes_len = [1 if rnd.random() < 0.25 else min(1024, 1 + int(rnd.expovariate(1 / 40.0))) for _ in range(n)]
    paths = {}
    for h in range(7):
        p = [v for v in range(n) if rnd.random() < 0.8]
        if h == 3:
            p = p[::-1]
        if h == 5:
            p = p + p[100:140]  # a path that visits nodes twice
        paths["GCF_%09d.%d" % (900 + 37 * h, 1 + h % 2)] = p
    paths["Z#last"] = [5, 4, 3]
    raw = bincode_graph(nodes_len, paths)
    with open(os.path.join(HERE, "graph_fixture.bin"), "wb") as f:
        f.write(raw)
    with open(os.path.join(HERE, "graph_fixture.bin.lz4"), "wb") as f:
        f.write(lz4_frame(raw))
    with open(os.path.join(HERE, "graph_fixture.bin.zst"), "wb") as f:
        f.write(zstd_frame(raw))
    with open(os.path.join(HERE, "graph_fixture.json"), "w") as f:
        json.dump({"nodes_len": nodes_len, "paths": paths}, f)
    print(len(raw), "bytes of bincode")


if __name__ == "__main__":
    main()
