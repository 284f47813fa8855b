"""Bit-parallel row formulation of the fill (prototype of nwb_batch_bp.cuh), checked against the oracle.

One row of the table is a handful of bit-vectors over the columns (bit i-1 = column i).  With
r(i,j) = score(i,j) + d(i+j) the differences u = r(i,j) - r(i-1,j) and v = r(i,j) - r(i,j-1) lie in
[0, M], M = 2d + m (nwb_fill_pk.cuh); per cell

    z = max(a, vL, uU),  u = z - vL,  v = z - uU          a = M on a match, N = 2d - k otherwise
    DIAG <=> z == a,  LEFT <=> u == 0,  UP <=> v == 0      (needleman-wunsch.c:485-503: every tie gets its arrow)

Along a row v(i) = max(y(i), v(i-1) - uU(i)) with y = max(a - uU, 0): level sets V_k = [v >= k] satisfy

    V_k(i) = S_k(i) | (V_k(i-1) & P(i)),   P = [uU == 0],
    S_k = [y >= k] | OR_{t>=1} ((V_{k+t} << 1) & [uU <= t])

and "seeds S propagate through runs of P" is one addition (Myers): V = S | (((S & P') + P') ^ P'), P' = P >> 1.
M additions per row, levels from M down to 1.
"""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def ge(planes, k, ones):
    """[x >= k] from binary planes (planes[0] = LSB)."""
    nb = len(planes)
    if k <= 0:
        return ones
    if k >= (1 << nb):
        return 0
    res = ones
    # evaluate from the LSB upwards: res = [low bits >= low bits of k]
    for t in range(nb):
        kb = (k >> t) & 1
        b = planes[t]
        res = (b & res) if kb else (b | res)
    return res


def to_binary(therm, nb):
    """binary planes of a thermometer code therm[1..M] (therm[k] = [x >= k])."""
    M = len(therm) - 1
    planes = []
    for t in range(nb):
        p = 0
        # bit t of x is set iff x in [2^t(2q+1), 2^t(2q+2)) for some q
        q = 0
        while (1 << t) * (2 * q + 1) <= M:
            lo = (1 << t) * (2 * q + 1)
            hi = (1 << t) * (2 * q + 2)
            p |= therm[lo] & ~(therm[hi] if hi <= M else 0)
            q += 1
        planes.append(p)
    return planes


def add_sub(v, uu, vl, nb, ones):
    """v + uu - vl mod 2^nb on binary planes."""
    # t = v - vl
    t = []
    borrow = 0
    for i in range(nb):
        t.append(v[i] ^ vl[i] ^ borrow)
        borrow = (~v[i] & (vl[i] | borrow) | (vl[i] & borrow)) & ones
    out = []
    carry = 0
    for i in range(nb):
        out.append(t[i] ^ uu[i] ^ carry)
        carry = (t[i] & uu[i]) | (carry & (t[i] ^ uu[i]))
    return out


def bp_fill(top, side, m, k, d):
    """Returns (codes[j][i] for j=1..B, i=1..A as list of lists, score)."""
    A, B = len(top), len(side)
    M, N = 2 * d + m, 2 * d - k
    assert 0 <= N <= M
    nb = max(1, M.bit_length())
    ones = (1 << A) - 1
    peq = {}
    for i, c in enumerate(top):
        peq[c] = peq.get(c, 0) | (1 << i)
    uu = [0] * nb  # u of the row above, binary planes
    rows = []
    for j in range(B):
        E = peq.get(side[j], 0)
        nE = ~E & ones
        UU = [None] + [ge(uu, t, ones) for t in range(1, M + 2)]  # UU[t] = [uU >= t]
        Y = [None] * (M + 1)
        for kk in range(1, M + 1):
            y = E & ~UU[M - kk + 1]
            if kk <= N:
                y |= nE & ~UU[N - kk + 1]
            Y[kk] = y & ones
        P = ~UU[1] & ones
        Pp = P >> 1
        V = [None] * (M + 2)
        Vs = [None] * (M + 2)
        V[M + 1] = 0
        for kk in range(M, 0, -1):
            S = Y[kk]
            for t in range(1, M - kk + 1):
                S |= Vs[kk + t] & ~UU[t + 1]
            S &= ones
            V[kk] = (S | (((S & Pp) + Pp) ^ Pp)) & ones
            Vs[kk] = (V[kk] << 1) & ones
        v = to_binary(V[:M + 1], nb)
        vl = to_binary(Vs[:M + 1], nb)
        un = add_sub(v, uu, vl, nb, ones)
        un = [x & ones for x in un]
        # z = v + uU
        z = []
        carry = 0
        for i in range(nb):
            z.append((v[i] ^ uu[i] ^ carry) & ones)
            carry = (v[i] & uu[i]) | (carry & (v[i] ^ uu[i]))
        # z == a: a = M where E else N
        eqM, eqN = ones, ones
        for i in range(nb):
            eqM &= z[i] if (M >> i) & 1 else ~z[i]
            eqN &= z[i] if (N >> i) & 1 else ~z[i]
        diag = ((E & eqM) | (nE & eqN)) & ones
        left = ones
        for i in range(nb):
            left &= ~un[i]
        left &= ones
        up = ~V[1] & ones
        rows.append([((diag >> i) & 1) | (((left >> i) & 1) << 1) | (((up >> i) & 1) << 2) for i in range(A)])
        uu = un
    # r(A,B) = sum_i u(i,B); score = r - d(A+B)
    r = 0
    for i in range(nb):
        r += bin(uu[i]).count("1") << i
    return rows, r - d * (A + B)


def main():
    import oracle
    rnd = random.Random(7)
    cases = 0
    for trial in range(400):
        alpha = rnd.choice([oracle.DNA, oracle.DNA, oracle.PROTEIN, "AB", "A"])
        A, B = rnd.randint(1, 70), rnd.randint(1, 70)
        if trial % 50 == 0:
            A, B = 256, 256
        top = "".join(rnd.choice(alpha) for _ in range(A))
        side = "".join(rnd.choice(alpha) for _ in range(B))
        while True:
            d = rnd.randint(0, 3)
            m = rnd.randint(0, 3)
            k = rnd.randint(-2, 2 * d)
            if 0 <= 2 * d - k <= 2 * d + m <= 7:
                break
        if trial % 3 == 0:
            m, k, d = 1, 1, 1
        o = oracle.fill(top, side, m, k, d, want_codes=True)
        rows, score = bp_fill(top, side, m, k, d)
        assert score == o.final_score, (trial, m, k, d, score, o.final_score)
        for j in range(1, B + 1):
            for i in range(1, A + 1):
                assert rows[j - 1][i - 1] == (int(o.codes[j, i]) & 7), (trial, m, k, d, i, j, rows[j - 1][i - 1], int(o.codes[j, i]))
        cases += 1
    print("bit-parallel rows match the oracle on", cases, "cases")


if __name__ == "__main__":
    main()
