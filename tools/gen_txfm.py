#!/usr/bin/env python3
"""Generate the 1-D AV1 integer transform kernels (DCT 4..32, ADST 8/16) as straight-line C.

The networks are described ONCE, as the forward butterfly stages (add/sub stages and
rotation stages with cospi[] weights, libaom `av1_fdct*/av1_fadst*` design).  The AV1
normative inverse is the same network run backwards with every stage matrix transposed
(every add/sub butterfly matrix is symmetric, every rotation is a 2x2 block), which this
script derives mechanically, so forward and inverse cannot drift apart.

Output: straight-line code on int32 locals, `half_btf` = round_shift(w0*a + w1*b, cos_bit)
evaluated in int64.  The same text is emitted for the C oracle (oracle/txfm1d_gen.h) and
for the CUDA device code (cavif_rs_b200/csrc/txfm1d_gen.h); both are pinned against
libaom 3.13.1's av1_fdct*/av1_idct*/av1_fadst*/av1_iadst* by tests/test_txfm.py.
"""
import os

def brev(bits, x):
    r = 0
    for i in range(bits):
        r |= ((x >> i) & 1) << (bits - 1 - i)
    return r

# A stage is a dict: out_wire -> list of (kind, coef, in_wire); kind 'one' (coef = +-1) or 'cos'
# (coef = +-k meaning +-cospi[k]).  Wires not mentioned pass through.

def add_stage(pairs):
    """pairs: list of (dst, [(sign, src), ...])"""
    return {d: [('one', s, w) for (s, w) in terms] for d, terms in pairs}

def rot_stage(rows):
    """rows: list of (dst, (k0, a), (k1, b)) -> dst = hb(sign*cospi[|k0|]*a + ...). k0/k1 are signed
    ints, with the convention that the sign is carried separately: give tuples (sign, k)."""
    out = {}
    for d, (c0, a), (c1, b) in rows:
        out[d] = [('cos', c0, a), ('cos', c1, b)]
    return out

def C(k): return (1, k)
def N(k): return (-1, k)

# ---- forward DCT odd parts, local wire j == global index M + j --------------------------------
def odd_stages(M):
    if M == 2:
        return [rot_stage([(0, (C(48), 0), (C(16), 1)), (1, (C(48), 1), (N(16), 0))])]
    if M == 4:
        return [
            rot_stage([(1, (N(32), 1), (C(32), 2)), (2, (C(32), 2), (C(32), 1))]),
            add_stage([(0, [(1, 0), (1, 1)]), (1, [(1, 0), (-1, 1)]), (2, [(1, 3), (-1, 2)]), (3, [(1, 3), (1, 2)])]),
            rot_stage([(0, (C(56), 0), (C(8), 3)), (1, (C(24), 1), (C(40), 2)),
                       (2, (C(24), 2), (N(40), 1)), (3, (C(56), 3), (N(8), 0))]),
        ]
    if M == 8:
        return [
            rot_stage([(2, (N(32), 2), (C(32), 5)), (3, (N(32), 3), (C(32), 4)),
                       (4, (C(32), 4), (C(32), 3)), (5, (C(32), 5), (C(32), 2))]),
            add_stage([(0, [(1, 0), (1, 3)]), (1, [(1, 1), (1, 2)]), (2, [(1, 1), (-1, 2)]), (3, [(1, 0), (-1, 3)]),
                       (4, [(1, 7), (-1, 4)]), (5, [(1, 6), (-1, 5)]), (6, [(1, 6), (1, 5)]), (7, [(1, 7), (1, 4)])]),
            rot_stage([(1, (N(16), 1), (C(48), 6)), (2, (N(48), 2), (N(16), 5)),
                       (5, (C(48), 5), (N(16), 2)), (6, (C(16), 6), (C(48), 1))]),
            add_stage([(0, [(1, 0), (1, 1)]), (1, [(1, 0), (-1, 1)]), (2, [(1, 3), (-1, 2)]), (3, [(1, 3), (1, 2)]),
                       (4, [(1, 4), (1, 5)]), (5, [(1, 4), (-1, 5)]), (6, [(1, 7), (-1, 6)]), (7, [(1, 7), (1, 6)])]),
            rot_stage([(0, (C(60), 0), (C(4), 7)), (1, (C(28), 1), (C(36), 6)), (2, (C(44), 2), (C(20), 5)),
                       (3, (C(12), 3), (C(52), 4)), (4, (C(12), 4), (N(52), 3)), (5, (C(44), 5), (N(20), 2)),
                       (6, (C(28), 6), (N(36), 1)), (7, (C(60), 7), (N(4), 0))]),
        ]
    if M == 16:
        return [
            rot_stage([(4, (N(32), 4), (C(32), 11)), (5, (N(32), 5), (C(32), 10)), (6, (N(32), 6), (C(32), 9)),
                       (7, (N(32), 7), (C(32), 8)), (8, (C(32), 8), (C(32), 7)), (9, (C(32), 9), (C(32), 6)),
                       (10, (C(32), 10), (C(32), 5)), (11, (C(32), 11), (C(32), 4))]),
            add_stage([(0, [(1, 0), (1, 7)]), (1, [(1, 1), (1, 6)]), (2, [(1, 2), (1, 5)]), (3, [(1, 3), (1, 4)]),
                       (4, [(1, 3), (-1, 4)]), (5, [(1, 2), (-1, 5)]), (6, [(1, 1), (-1, 6)]), (7, [(1, 0), (-1, 7)]),
                       (8, [(1, 15), (-1, 8)]), (9, [(1, 14), (-1, 9)]), (10, [(1, 13), (-1, 10)]), (11, [(1, 12), (-1, 11)]),
                       (12, [(1, 12), (1, 11)]), (13, [(1, 13), (1, 10)]), (14, [(1, 14), (1, 9)]), (15, [(1, 15), (1, 8)])]),
            rot_stage([(2, (N(16), 2), (C(48), 13)), (3, (N(16), 3), (C(48), 12)),
                       (4, (N(48), 4), (N(16), 11)), (5, (N(48), 5), (N(16), 10)),
                       (10, (C(48), 10), (N(16), 5)), (11, (C(48), 11), (N(16), 4)),
                       (12, (C(16), 12), (C(48), 3)), (13, (C(16), 13), (C(48), 2))]),
            add_stage([(0, [(1, 0), (1, 3)]), (1, [(1, 1), (1, 2)]), (2, [(1, 1), (-1, 2)]), (3, [(1, 0), (-1, 3)]),
                       (4, [(1, 7), (-1, 4)]), (5, [(1, 6), (-1, 5)]), (6, [(1, 6), (1, 5)]), (7, [(1, 7), (1, 4)]),
                       (8, [(1, 8), (1, 11)]), (9, [(1, 9), (1, 10)]), (10, [(1, 9), (-1, 10)]), (11, [(1, 8), (-1, 11)]),
                       (12, [(1, 15), (-1, 12)]), (13, [(1, 14), (-1, 13)]), (14, [(1, 14), (1, 13)]), (15, [(1, 15), (1, 12)])]),
            rot_stage([(1, (N(8), 1), (C(56), 14)), (2, (N(56), 2), (N(8), 13)),
                       (5, (N(40), 5), (C(24), 10)), (6, (N(24), 6), (N(40), 9)),
                       (9, (C(24), 9), (N(40), 6)), (10, (C(40), 10), (C(24), 5)),
                       (13, (C(56), 13), (N(8), 2)), (14, (C(8), 14), (C(56), 1))]),
            add_stage([(0, [(1, 0), (1, 1)]), (1, [(1, 0), (-1, 1)]), (2, [(1, 3), (-1, 2)]), (3, [(1, 3), (1, 2)]),
                       (4, [(1, 4), (1, 5)]), (5, [(1, 4), (-1, 5)]), (6, [(1, 7), (-1, 6)]), (7, [(1, 7), (1, 6)]),
                       (8, [(1, 8), (1, 9)]), (9, [(1, 8), (-1, 9)]), (10, [(1, 11), (-1, 10)]), (11, [(1, 11), (1, 10)]),
                       (12, [(1, 12), (1, 13)]), (13, [(1, 12), (-1, 13)]), (14, [(1, 15), (-1, 14)]), (15, [(1, 15), (1, 14)])]),
            rot_stage([(0, (C(62), 0), (C(2), 15)), (1, (C(30), 1), (C(34), 14)), (2, (C(46), 2), (C(18), 13)),
                       (3, (C(14), 3), (C(50), 12)), (4, (C(54), 4), (C(10), 11)), (5, (C(22), 5), (C(42), 10)),
                       (6, (C(38), 6), (C(26), 9)), (7, (C(6), 7), (C(58), 8)), (8, (C(6), 8), (N(58), 7)),
                       (9, (C(38), 9), (N(26), 6)), (10, (C(22), 10), (N(42), 5)), (11, (C(54), 11), (N(10), 4)),
                       (12, (C(14), 12), (N(50), 3)), (13, (C(46), 13), (N(18), 2)), (14, (C(30), 14), (N(34), 1)),
                       (15, (C(62), 15), (N(2), 0))]),
        ]
    raise ValueError(M)

def dct_stages(n):
    """Full forward DCT network on n wires; input natural order, output natural order."""
    if n == 2:
        return [rot_stage([(0, (C(32), 0), (C(32), 1)), (1, (N(32), 1), (C(32), 0))])]
    h = n // 2
    st = []
    # split: e[i] = x[i] + x[n-1-i] (wire i), o[j] = x[h-1-j] - x[h+j] (wire h+j)
    st.append(add_stage([(i, [(1, i), (1, n - 1 - i)]) for i in range(h)] +
                        [(h + j, [(1, h - 1 - j), (-1, h + j)]) for j in range(h)]))
    # even part on wires 0..h-1 (recursive), odd part on wires h..n-1
    st += dct_stages(h)
    for s in odd_stages(h):
        st.append({h + d: [(k, c, h + w) for (k, c, w) in terms] for d, terms in s.items()})
    # output permutation: E[k] -> X[2k], O-wire j -> X[2*brev(j)+1]
    bits = h.bit_length() - 1
    perm = {}
    for k in range(h):
        perm[2 * k] = [('one', 1, k)]
    for j in range(h):
        perm[2 * brev(bits, j) + 1] = [('one', 1, h + j)]
    st.append(perm)
    return st

def adst_stages(n):
    if n == 8:
        return [
            {0: [('one', 1, 0)], 1: [('one', -1, 7)], 2: [('one', -1, 3)], 3: [('one', 1, 4)],
             4: [('one', -1, 1)], 5: [('one', 1, 6)], 6: [('one', 1, 2)], 7: [('one', -1, 5)]},
            rot_stage([(2, (C(32), 2), (C(32), 3)), (3, (C(32), 2), (N(32), 3)),
                       (6, (C(32), 6), (C(32), 7)), (7, (C(32), 6), (N(32), 7))]),
            add_stage([(0, [(1, 0), (1, 2)]), (1, [(1, 1), (1, 3)]), (2, [(1, 0), (-1, 2)]), (3, [(1, 1), (-1, 3)]),
                       (4, [(1, 4), (1, 6)]), (5, [(1, 5), (1, 7)]), (6, [(1, 4), (-1, 6)]), (7, [(1, 5), (-1, 7)])]),
            rot_stage([(4, (C(16), 4), (C(48), 5)), (5, (C(48), 4), (N(16), 5)),
                       (6, (N(48), 6), (C(16), 7)), (7, (C(16), 6), (C(48), 7))]),
            add_stage([(0, [(1, 0), (1, 4)]), (1, [(1, 1), (1, 5)]), (2, [(1, 2), (1, 6)]), (3, [(1, 3), (1, 7)]),
                       (4, [(1, 0), (-1, 4)]), (5, [(1, 1), (-1, 5)]), (6, [(1, 2), (-1, 6)]), (7, [(1, 3), (-1, 7)])]),
            rot_stage([(0, (C(4), 0), (C(60), 1)), (1, (C(60), 0), (N(4), 1)), (2, (C(20), 2), (C(44), 3)),
                       (3, (C(44), 2), (N(20), 3)), (4, (C(36), 4), (C(28), 5)), (5, (C(28), 4), (N(36), 5)),
                       (6, (C(52), 6), (C(12), 7)), (7, (C(12), 6), (N(52), 7))]),
            {0: [('one', 1, 1)], 1: [('one', 1, 6)], 2: [('one', 1, 3)], 3: [('one', 1, 4)],
             4: [('one', 1, 5)], 5: [('one', 1, 2)], 6: [('one', 1, 7)], 7: [('one', 1, 0)]},
        ]
    if n == 16:
        s1 = {0: (1, 0), 1: (-1, 15), 2: (-1, 7), 3: (1, 8), 4: (-1, 3), 5: (1, 12), 6: (1, 4), 7: (-1, 11),
              8: (-1, 1), 9: (1, 14), 10: (1, 6), 11: (-1, 9), 12: (1, 2), 13: (-1, 13), 14: (-1, 5), 15: (1, 10)}
        outp = {0: 1, 1: 14, 2: 3, 3: 12, 4: 5, 5: 10, 6: 7, 7: 8, 8: 9, 9: 6, 10: 11, 11: 4, 12: 13, 13: 2, 14: 15, 15: 0}
        def pm(base, step):  # butterflies a+b / a-b between wire i and i+step in groups
            rows = []
            for g in range(0, 16, 2 * step):
                for i in range(step):
                    a, b = g + i, g + i + step
                    rows.append((a, [(1, a), (1, b)]))
                    rows.append((b, [(1, a), (-1, b)]))
            return add_stage(rows)
        return [
            {d: [('one', s, w)] for d, (s, w) in s1.items()},
            rot_stage([(b, (C(32), b), (C(32), b + 1)) for b in (2, 6, 10, 14)] +
                      [(b + 1, (C(32), b), (N(32), b + 1)) for b in (2, 6, 10, 14)]),
            pm(0, 2),
            rot_stage([(4, (C(16), 4), (C(48), 5)), (5, (C(48), 4), (N(16), 5)), (6, (N(48), 6), (C(16), 7)),
                       (7, (C(16), 6), (C(48), 7)), (12, (C(16), 12), (C(48), 13)), (13, (C(48), 12), (N(16), 13)),
                       (14, (N(48), 14), (C(16), 15)), (15, (C(16), 14), (C(48), 15))]),
            pm(0, 4),
            rot_stage([(8, (C(8), 8), (C(56), 9)), (9, (C(56), 8), (N(8), 9)), (10, (C(40), 10), (C(24), 11)),
                       (11, (C(24), 10), (N(40), 11)), (12, (N(56), 12), (C(8), 13)), (13, (C(8), 12), (C(56), 13)),
                       (14, (N(24), 14), (C(40), 15)), (15, (C(40), 14), (C(24), 15))]),
            pm(0, 8),
            rot_stage([(0, (C(2), 0), (C(62), 1)), (1, (C(62), 0), (N(2), 1)), (2, (C(10), 2), (C(54), 3)),
                       (3, (C(54), 2), (N(10), 3)), (4, (C(18), 4), (C(46), 5)), (5, (C(46), 4), (N(18), 5)),
                       (6, (C(26), 6), (C(38), 7)), (7, (C(38), 6), (N(26), 7)), (8, (C(34), 8), (C(30), 9)),
                       (9, (C(30), 8), (N(34), 9)), (10, (C(42), 10), (C(22), 11)), (11, (C(22), 10), (N(42), 11)),
                       (12, (C(50), 12), (C(14), 13)), (13, (C(14), 12), (N(50), 13)), (14, (C(58), 14), (C(6), 15)),
                       (15, (C(6), 14), (N(58), 15))]),
            {d: [('one', 1, w)] for d, w in outp.items()},
        ]
    raise ValueError(n)

def transpose(stage, n):
    full = {}
    for d in range(n):
        full[d] = stage.get(d, [('one', 1, d)])
    out = {w: [] for w in range(n)}
    for d, terms in full.items():
        for (kind, c, w) in terms:
            out[w].append((kind, c, d))
    return out

def emit_fn(name, n, stages, f, wide=True):
    """wide=True: half_btf evaluated in int64 (always exact).  wide=False: int32 multiply-accumulate, exact as long
    as |w0*a + w1*b| < 2^31 - which holds for 8-bit video (see DESIGN.md 'int32 butterflies')."""
    f.write("TXFM_FN void %s(const int32_t *in, int32_t *out, const int32_t *cospi, int bit) {\n" % name)
    cur = ["in[%d]" % i for i in range(n)]
    for si, st in enumerate(stages):
        nxt = list(cur)
        decl = []
        for d in range(n):
            if d not in st:
                continue
            terms = st[d]
            if len(terms) == 1 and terms[0][0] == 'one':
                _, s, w = terms[0]
                if s == 1:
                    nxt[d] = cur[w]
                    continue
                expr = "-%s" % cur[w]
            elif all(t[0] == 'one' for t in terms):
                expr = ""
                for (_, s, w) in terms:
                    expr += (" + " if s > 0 else " - ") + cur[w]
                expr = expr[3:] if expr.startswith(" + ") else "-" + expr[3:]
            else:
                assert all(t[0] == 'cos' for t in terms) and len(terms) == 2, (name, si, d, terms)
                parts = []
                for (_, (s, k), w) in terms:
                    parts.append("%s%scospi[%d] * %s" % ("-" if s < 0 else "", "(int64_t)" if wide else "", k, cur[w]))
                expr = "(int32_t)((%s + %s + rnd) >> bit)" % (parts[0], parts[1])
            v = "s%d_%d" % (si, d)
            decl.append("  const int32_t %s = %s;\n" % (v, expr))
            nxt[d] = v
        if any("rnd" in x for x in decl) and "rnd_declared" not in emit_fn.__dict__.get(name, ""):
            pass
        f.writelines(decl)
        cur = nxt
    for d in range(n):
        f.write("  out[%d] = %s;\n" % (d, cur[d]))
    f.write("}\n\n")

def generate(path):
    with open(path, "w") as f:
        f.write("// GENERATED by tools/gen_txfm.py - AV1 1-D DCT/ADST butterfly networks. Do not edit.\n")
        f.write("// in/out must not alias.  cospi = av1_cospi_arr[bit-10], half_btf in int64.\n")
        f.write("#ifndef TXFM1D_GEN_H_\n#define TXFM1D_GEN_H_\n#include <stdint.h>\n")
        f.write("#ifndef TXFM_FN\n#define TXFM_FN static inline\n#endif\n\n")
        for n in (4, 8, 16, 32):
            st = dct_stages(n)
            body_f = st
            body_i = [transpose(s, n) for s in reversed(st)]
            for nm, body in (("gen_fdct%d" % n, body_f), ("gen_idct%d" % n, body_i)):
                import io
                for wide in (True, False):
                    if not wide and n == 32:
                        continue
                    buf = io.StringIO()
                    emit_fn(nm + ("" if wide else "_w32"), n, body, buf, wide)
                    ty = "int64_t" if wide else "int32_t"
                    txt = buf.getvalue().replace("{\n", "{\n  const %s rnd = (%s)1 << (bit - 1);\n" % (ty, ty), 1)
                    f.write(txt)
        for n in (8, 16):
            st = adst_stages(n)
            body_i = [transpose(s, n) for s in reversed(st)]
            for nm, body in (("gen_fadst%d" % n, st), ("gen_iadst%d" % n, body_i)):
                import io
                for wide in (True, False):
                    buf = io.StringIO()
                    emit_fn(nm + ("" if wide else "_w32"), n, body, buf, wide)
                    ty = "int64_t" if wide else "int32_t"
                    txt = buf.getvalue().replace("{\n", "{\n  const %s rnd = (%s)1 << (bit - 1);\n" % (ty, ty), 1)
                    f.write(txt)
        f.write("#endif  // TXFM1D_GEN_H_\n")

if __name__ == "__main__":
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    for p in ("oracle/txfm1d_gen.h", "cavif_rs_b200/csrc/txfm1d_gen.h"):
        generate(os.path.join(root, p))
        print("wrote", p)
