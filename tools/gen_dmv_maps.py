#!/usr/bin/env python
"""tools/gen_dmv_maps.py -- lane mappings of the gather-schedule DMV kernel (vlgae_b200/csrc/dmv_gather.cu).

Every width step of the chart sweep hands the (span, split point) rectangle of that width to the lanes of a CTA:
G = 2^lg lanes per span, each lane a contiguous chunk of c split points, lanes ordered span-major or chunk-major.
The kernel is bound by shared-memory wavefronts, and which (G, c, order) is conflict-free and
fills its half-warps depends on the width, the sentence length and the row stride.  This script counts the wavefronts
of every candidate with a bank model of the B200 shared memory (32 banks x 4 B; a 64-bit access is served per half-warp)
and writes the best choice per (pass, positions, width) to vlgae_b200/csrc/dmv_maps.inc.  Runs on the CPU.

    python tools/gen_dmv_maps.py            # rewrites dmv_maps.inc
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def wavefronts(acc, width):
    """acc: list of (lane, byte address) of the active lanes of one warp-wide access."""
    if not acc:
        return 0
    tot = 0
    if width == 8:
        for hw in (0, 1):
            d = {}
            for lane, a in acc:
                if lane // 16 == hw:
                    d.setdefault((a // 8) % 16, set()).add(a)
            if d:
                tot += max(len(v) for v in d.values())
    else:
        d = {}
        for lane, a in acc:
            d.setdefault((a // 4) % 32, set()).add(a)
        tot += max(len(v) for v in d.values())
    return tot


# streams: (bytes per lane, address(i, j, a), accesses per term (2 = read-modify-write))
def streams_inside(S):
    return [
        (8, lambda i, j, a: 8 * (i * S + i + a + 1), 1),   # CR(i, i+a)       C[i][i+a+1]
        (8, lambda i, j, a: 8 * (j * S + i + a + 1), 1),   # CL(j, i+a+1)     C[j][i+a+1]
        (8, lambda i, j, a: 8 * (j * S + i + a + 1), 1),   # IL(j, i+a+1)     I[j][i+a+1]
        (8, lambda i, j, a: 8 * (i * S + i + a + 1), 1),   # IR(i, i+a+1)     I[i][i+a+1]
        (4, lambda i, j, a: 4 * (i * S + i + a + 2), 1),   # CLn(i+a+1, i)    Ct[i][i+a+2]
        (4, lambda i, j, a: 4 * (j * S + i + a + 1), 1),   # CRn(i+a+1, j)    Ct[j][i+a+1]
    ]


def streams_aprime(S):
    return [
        (4, lambda i, j, a: 8 * ((i + a) * S + i), 1),              # CLn(i+a, i)   C[i+a][i].x
        (8, lambda i, j, a: 8 * (j * S + i + a), 1),                # IL(j, i+a)
        (8, lambda i, j, a: 8 * (j * S + i + a), 2),                # gIL(j, i+a)   read-modify-write
        (4, lambda i, j, a: 8 * ((i + a) * S + i), 2),              # gCLn(i+a, i)  read-modify-write
        (4, lambda i, j, a: 8 * ((i + 1 + a) * S + j + 1) + 4, 1),  # CRn(i+1+a, j) C[i+1+a][j+1].y
        (8, lambda i, j, a: 8 * (i * S + i + 1 + a), 1),            # IR(i, i+1+a)
        (8, lambda i, j, a: 8 * (i * S + i + 1 + a), 2),            # gIR
        (4, lambda i, j, a: 8 * ((i + 1 + a) * S + j + 1) + 4, 2),  # gCRn
    ]


def streams_bprime(S):
    return [
        (8, lambda i, j, a: 8 * (i * S + i + a + 1), 1),
        (8, lambda i, j, a: 8 * (j * S + i + a + 1), 1),
        (8, lambda i, j, a: 8 * (i * S + i + a + 1), 2),
        (8, lambda i, j, a: 8 * (j * S + i + a + 1), 2),
    ]


def lanes_of(lg, c, qmajor, base, n, T):
    G = 1 << lg
    spw = 32 >> lg
    out = []
    for lane in range(32):
        if qmajor:
            q, il = lane // spw, lane % spw
        else:
            q, il = lane % G, lane // G
        i = base + il
        if i < n and q * c < T:
            out.append((lane, i, range(q * c, min(T, q * c + c))))
    return out


def cost_of(Nb, w, T, lg, c, qmajor, streams, term_clk, reduce_vals, max_warps):
    """(shared-memory wavefronts, dependent-chain estimate in clocks) of one width step under a mapping."""
    n = Nb - w
    spw = 32 >> lg
    wf = 0
    nwarps = 0
    for base in range(0, n, spw):
        lanes = lanes_of(lg, c, qmajor, base, n, T)
        if not lanes:
            continue
        nwarps += 1
        kmax = max(len(r) for _, _, r in lanes)
        for k in range(kmax):
            for width, f, mult in streams:
                acc = [(lane, f(i, i + w, r[k])) for lane, i, r in lanes if k < len(r)]
                wf += mult * wavefronts(acc, width)
    rounds = -(-nwarps // max_warps)
    # one round of a warp (measured with tools/probes/lds_batch_probe.cu: a lone warp spends ~7 clocks per shared-memory
    # instruction whatever the batching): set-up, c split points, xor-shuffle rounds, finalisation
    chain = rounds * (60 + term_clk * c + (45 * lg if reduce_vals else 0) + (110 if reduce_vals else 30))
    return wf, chain


def best_map(Nb, w, T, streams, term_clk, reduce_vals, max_warps=4):
    best = None
    for lg in range(0, 6):
        G = 1 << lg
        if lg > 0 and G > max(T, 1):
            continue
        c0 = max(1, -(-T // G))
        for c in range(c0, min(63, c0 + 3) + 1):
            if lg > 0 and (G - 1) * c >= T:
                continue  # a whole lane group would idle: a smaller G does the same
            for qmajor in (0, 1):
                wf, chain = cost_of(Nb, w, T, lg, c, qmajor, streams, term_clk, reduce_vals, max_warps)
                # With 4 resident sentences per SM the kernel is bound by the dependent chain of a width step first and
                # by shared-memory wavefronts (1 per clock per SM) second: 4 CTAs overlap their chains.
                cost = chain / 4.0 + wf
                if best is None or cost < best[0]:
                    best = (cost, lg, c, qmajor, wf)
    return best


def main():
    kinds = [("inside / Viterbi (fused steps 1-4, w - 1 split points)", streams_inside, lambda w: w - 1, 41, 6),
             ("reverse sweep, complete parents (w split points)", streams_aprime, lambda w: w, 80, 0),
             ("reverse sweep, incomplete parents (w split points)", streams_bprime, lambda w: w, 40, 0)]
    lines = ["// dmv_maps.inc -- generated by tools/gen_dmv_maps.py (bank model of the shared memory); do not edit.",
             "// entry = lg | c << 3 | chunk_major << 9 for (CTA size: 2, 4 or 8 warps; pass; positions Nb; width w), one table per row stride S"]
    for S in (26, 34, 42):
        CAP = S - 1
        lines.append("static __device__ const unsigned short g_dmv_map_%d[3][3][%d][%d] = {" % (S, CAP + 1, CAP + 1))
        tot_wf = [0, 0, 0]
        ideal = [0.0, 0.0, 0.0]
        for max_warps in (2, 4, 8):
            lines.append(" {  // CTAs of %d warps" % max_warps)
            for kidx, (name, sf, tf, ipt, rv) in enumerate(kinds):
                lines.append("  {  // " + name)
                streams = sf(S)
                for Nb in range(CAP + 1):
                    row = []
                    for w in range(CAP + 1):
                        T = tf(w)
                        if w < 1 or w >= Nb or T <= 0:
                            row.append(0)
                            continue
                        _, lg, c, qm, wf = best_map(Nb, w, T, streams, ipt, rv, max_warps)
                        row.append(lg | (c << 3) | (qm << 9))
                        if Nb == CAP and max_warps == 4:
                            tot_wf[kidx] += wf
                            ideal[kidx] += (Nb - w) * T * sum(wd * m for wd, _, m in streams) / 128.0
                    lines.append("    {" + ", ".join(str(v) for v in row) + "},")
                lines.append("  },")
            lines.append(" },")
        lines.append("};")
        for kidx, (name, *_rest) in enumerate(kinds):
            print(f"S = {S}: {name}: {tot_wf[kidx]} wavefronts at Nb = {CAP} (ideal {ideal[kidx]:.0f}, efficiency {ideal[kidx] / max(tot_wf[kidx], 1):.2f})", flush=True)
    with open(os.path.join(ROOT, "vlgae_b200", "csrc", "dmv_maps.inc"), "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
