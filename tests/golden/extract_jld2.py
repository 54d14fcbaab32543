#!/usr/bin/env python3
"""Extract the reference's golden objective/gradient vectors into JSON fixtures.

Reads  /root/reference/test/reference_solutions/*.jld2  (JLD2 = HDF5 superblock v2 at
byte 512, v2 object headers, compact float64 datasets) and the pcof input vectors
/root/reference/test/cases/*.dat, and writes tests/golden/<case>.json.

The JLD2 files were written by the reference's own regeneration script
(/root/reference/test/cases/refSol.jl:1-41 via test/evalGrad.jl:31) and are what
test/runtests.jl:30-54 compares against at rtol 1e-10 / atol 1e-14.

This script only runs in the build container (it needs /root/reference); the JSON it
emits is committed so that nothing at test time reads /root/reference.
No h5py in this image -> a ~60 line pure-Python walker of the object headers.
"""
import json, os, struct, sys

REF = "/root/reference/test"
BASE = 512  # JLD2 puts the HDF5 superblock after a 512-byte text header


def _parse_ohdr(d, addr):
    """Parse one v2 object header at absolute offset addr -> list of (type, body bytes)."""
    assert d[addr:addr + 4] == b"OHDR", (addr, d[addr:addr + 4])
    ver, flags = d[addr + 4], d[addr + 5]
    assert ver == 2
    p = addr + 6
    if flags & 0x20:
        p += 16  # times
    if flags & 0x10:
        p += 4  # max compact / min dense
    szlen = 1 << (flags & 3)
    chunk = int.from_bytes(d[p:p + szlen], "little")
    p += szlen
    end = p + chunk
    msgs = []
    while p + 4 <= end:
        mtype = d[p]
        msize = int.from_bytes(d[p + 1:p + 3], "little")
        p += 4
        if flags & 0x04:
            p += 2  # creation order
        msgs.append((mtype, d[p:p + msize]))
        p += msize
    return msgs


def _links(d, root):
    out = {}
    for mtype, body in _parse_ohdr(d, root):
        if mtype != 6:
            continue
        ver, lflags = body[0], body[1]
        q = 2
        if lflags & 0x08:
            q += 1  # link type
        if lflags & 0x04:
            q += 8  # creation order
        if lflags & 0x10:
            q += 1  # charset
        nlen_sz = 1 << (lflags & 3)
        nlen = int.from_bytes(body[q:q + nlen_sz], "little")
        q += nlen_sz
        name = body[q:q + nlen].decode()
        q += nlen
        out[name] = int.from_bytes(body[q:q + 8], "little") + BASE
    return out


def _dataset(d, addr):
    dims, data = [], None
    for mtype, body in _parse_ohdr(d, addr):
        if mtype == 1:  # dataspace v2
            rank = body[1]
            dims = [int.from_bytes(body[4 + 8 * i:12 + 8 * i], "little") for i in range(rank)]
        elif mtype == 3:
            assert body[0] & 0x0F == 1 and int.from_bytes(body[4:8], "little") == 8, "expect float64"
        elif mtype == 8:
            assert body[0] in (3, 4) and body[1] == 0, "expect compact layout"
            size = int.from_bytes(body[2:4], "little")
            data = list(struct.unpack("<%dd" % (size // 8), body[4:4 + size]))
    return dims, data


def read_jld2(path):
    d = open(path, "rb").read()
    assert d[BASE:BASE + 8] == b"\x89HDF\r\n\x1a\n" and d[BASE + 8] == 2
    root = int.from_bytes(d[BASE + 36:BASE + 44], "little") + BASE
    return {k: _dataset(d, a) for k, a in _links(d, root).items()}


def read_dat(path):
    return [float(x) for x in open(path).read().split()]


CASES = ["rabi", "swap02", "flux", "cnot2", "cnot3", "cnot2-leakieq", "cnot2-jacobi"]


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    for case in CASES:
        g = read_jld2(f"{REF}/reference_solutions/{case}-ref.jld2")
        rec = {"case": case,
               "source": f"test/reference_solutions/{case}-ref.jld2 (+ test/cases/{case}.dat)",
               "obj0": g["obj0"][1], "grad0": g["grad0"][1]}
        dat = f"{REF}/cases/{case}.dat"
        if case != "rabi":  # rabi starts from the analytic pcof (test/cases/rabi-setup.jl:150-157)
            rec["pcof0"] = read_dat(dat)
        with open(os.path.join(here, f"{case}.json"), "w") as f:
            json.dump(rec, f, indent=0)
        print(case, "obj0", rec["obj0"], "len(grad0)", len(rec["grad0"]), "len(pcof0)", len(rec.get("pcof0", [])))
    e = read_jld2(f"{REF}/reference_solutions/err-mat-ref.jld2")
    (name, (dims, data)), = e.items()
    with open(os.path.join(here, "err-mat.json"), "w") as f:
        json.dump({"name": name, "hdf5_dims": dims, "data": data,
                   "source": "test/reference_solutions/err-mat-ref.jld2"}, f)
    print("err-mat", name, dims, data[:2])
    # optimised pulses shipped with the reference's examples (examples/drives/*.jld2, key "pcof"): the EXAMPLE
    # configurations (examples/cnot2-setup.jl at T = 50, examples/rabi-setup.jl) must turn them into a high-fidelity gate
    drives = {}
    for f, cfgname in (("cnot2-pcof-opt-t50", "cnot2"), ("cnot2-pcof-opt-t100", "cnot2-T100"), ("cnot2-pcof-opt-t200", "cnot2-T200"),
                       ("rabi-pcof-opt-t100", "rabi"), ("cnot3-pcof-opt", "cnot3-Nfreq3")):
        g = read_jld2(f"/root/reference/examples/drives/{f}.jld2")
        drives[cfgname] = {"source": f"examples/drives/{f}.jld2", "pcof": g["pcof"][1]}
        print("drive", f, len(g["pcof"][1]))
    with open(os.path.join(here, "drives.json"), "w") as f:
        json.dump(drives, f, indent=0)


if __name__ == "__main__":
    sys.exit(main())
