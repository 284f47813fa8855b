"""Prototype of the per-lane sparse backward count (nwb_batch_lcount.cuh): one pair per thread, a window of 32
columns that follows the band of cells on optimal paths (shifts one column left per row by construction,
re-centred by +-SHIFT columns when the live cells reach a guard zone).  Returns (count mod 2^64, bailed).

    P(A,B) = 1,  P(i,j) = [DIAG(i+1,j+1)] P(i+1,j+1) + [LEFT(i+1,j)] P(i+1,j) + [UP(i,j+1)] P(i,j+1)
    count = flow that reaches row 0 / column 0 (computation.c:97-124: one forced arrow each)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

W = 32
MASK = (1 << 64) - 1


def lcount(codes, A, B, guard=8, shift=8, start=23):
    """codes[j][i] (1-based interior, border row/col ignored): DIAG=1, LEFT=2, UP=4."""
    def nib(col, j):
        if col == 0 and j >= 1:
            return 4            # the border column: forced UP
        if col < 1 or col > A or j < 1:
            return 0
        return int(codes[j][col]) & 7
    base = A - start
    inc = [0] * W
    inc[start] = 1
    for j in range(B, 0, -1):
        # re-centring decision for the NEXT row, from the live cells of this row before its update
        left = any(inc[k] for k in range(guard))
        right = any(inc[k] for k in range(W - guard, W))
        if left and right:
            return None, True
        sh = -shift if left else (shift if right else 0)
        x = [nib(base + k, j) for k in range(W)]
        d = 0
        dgp = 0
        new = [0] * W
        for k in range(W - 1, -1, -1):
            P = (inc[k] + d) & MASK
            d = P if x[k] & 2 else 0
            up = P if x[k] & 4 else 0
            dg = P if x[k] & 1 else 0
            if k == W - 1:
                if up:
                    return None, True
            else:
                new[k + 1] = (up + dgp) & MASK
            dgp = dg
        new[0] = dgp
        if d:
            return None, True
        # window of row j-1: base - 1 (built in), then the shift
        base -= 1
        if sh < 0:      # window moves left by `shift`: cells move right in the window
            if any(new[W - shift:]):
                return None, True
            new = [0] * shift + new[:W - shift]
            base -= shift
        elif sh > 0:
            if any(new[:shift]):
                return None, True
            new = new[shift:] + [0] * shift
            base += shift
        inc = new
    return sum(inc) & MASK, False


def main():
    import oracle
    import random
    rnd = random.Random(1)
    for guard, shift in ((8, 8), (6, 6), (4, 4), (6, 4), (8, 4)):
        bails = 0
        n = 0
        for p in range(300):
            t, s = oracle.generate_pair(0x5EED4000 + 2 * p, 256, 256)
            o = oracle.fill(t, s, 1, 1, 1, want_codes=True)
            c, b = lcount(o.codes, 256, 256, guard, shift)
            n += 1
            if b:
                bails += 1
            else:
                assert c == o.count, (p, c, o.count)
        print(f"guard {guard} shift {shift}: bailed {bails}/{n}")
    # ragged shapes and other alphabets: counts must match whenever the sweep does not give up
    ok = bail = 0
    for trial in range(300):
        a, b = rnd.randint(1, 300), rnd.randint(1, 300)
        alpha = rnd.choice(["ACGT", "AC", "A", oracle.PROTEIN])
        t = "".join(rnd.choice(alpha) for _ in range(a))
        s = "".join(rnd.choice(alpha) for _ in range(b))
        m, k, d = rnd.choice([(1, 1, 1), (2, 1, 2), (0, 0, 0), (1, 0, 1)])
        o = oracle.fill(t, s, m, k, d, want_codes=True)
        c, bl = lcount(o.codes, a, b)
        if bl:
            bail += 1
        else:
            assert c == o.count, (trial, a, b, c, o.count)
            ok += 1
    print("ragged:", ok, "ok,", bail, "gave up")


if __name__ == "__main__":
    main()
