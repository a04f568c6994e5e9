"""Concurrency / determinism hunt on the GPU box: T threads x M rounds of the RDO test seam on different frames; every
output array of every call is compared with the serial result of the same frame.  Prints which arrays ever differed."""
import os, sys, threading, collections, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import harness as H
import cpu_encode as CE

L = H.hostlib()


def device_rdo(f, src):
    out = H.Records(f)
    out.deblocked = [np.zeros_like(r) for r in out.rec]
    out.final = [np.zeros_like(r) for r in out.rec]
    out.lr_out = [np.zeros_like(r) for r in out.rec]
    src3 = list(src) + [src[0]] * (3 - len(src))
    rc = L.b200av1_rdo_run(ctypes.byref(f), H.ptr3(src3), out.mi.ctypes.data, H.ptr3(out.eob), H.ptr3(out.coef), H.ptr3(out.rec),
                           H.ptr3(out.deblocked), H.ptr3(out.final), out.cdef_idx.ctypes.data, H.ptrn(out.dec), H.ptrn(out.satd),
                           H.ptr3(out.lr), H.ptr3(out.lr_out))
    assert rc == 0
    return out


def fields(o, planes):
    d = {"mi": o.mi, "cdef_idx": o.cdef_idx}
    for lv in range(H.LEVELS):
        d["dec%d" % lv] = o.dec[lv]; d["satd%d" % lv] = o.satd[lv]
    for p in range(planes):
        for k in ("eob", "coef", "rec", "deblocked", "final", "lr", "lr_out"):
            d["%s%d" % (k, p)] = getattr(o, k)[p]
    return d


def main(threads=8, rounds=12):
    frames = []
    for i in range(threads):
        w, h = 160 + 16 * i, 96 + 8 * i
        mono = i % 3 == 0
        img = H.synth_image(w, h, 20 + i, alpha=True)
        pl, alpha, _ = CE.convert(img, 10)
        planes = [alpha] if mono else pl
        f = CE.configure(w, h, 10, len(planes), 125, 6)
        f.full_range = 1
        frames.append((f, H.pad_planes(planes, f), len(planes)))
    serial = [fields(device_rdo(f, src), n) for f, src, n in frames]
    bad = collections.Counter()
    def work(i):
        f, src, n = frames[i]
        for _ in range(rounds):
            got = fields(device_rdo(f, src), n)
            for k, v in got.items():
                if not np.array_equal(v, serial[i][k]):
                    bad[(i, k)] += 1
    th = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    for t in th: t.start()
    for t in th: t.join()
    print("serial_recon" if os.environ.get("B200AVIF_DEBUG_SERIAL_RECON") else "wavefront", "mismatches:", dict(bad) if bad else "none")


if __name__ == "__main__":
    main()
