"""CPU model of the small-angle inner solve of k_jacobi_blk (xtb_b200/csrc/kernels_linalg.cuh, jb_sweeps2, "simultaneous
rotations"): the lane mapping of the kernel restated in numpy -- which lane evaluates which rotation, how J is accumulated in
registers with one shuffle per round, the m8n8k4 fragment indices of the two 8 x 8 x 8 products -- and the block Jacobi solver
around it, with and without the simultaneous form.  What the eigensolver replaces: dsyevd in
src/mctc/lapack/eigensolve.F90:218-246 (its results are pinned through the energies, see tests/test_gpu_parity.py)."""
import numpy as np

KJS = 12                      # row stride of the per-warp 8 x 8 scratch matrices (kJs)
LANES = range(32)


def rotation(app, aqq, apq):
    """(c, s) of jb_rotation: |theta| <= pi/4, eliminates M[p][q] under col p <- c col p - s col q, col q <- s col p + c col q."""
    if apq * apq > 4.93e-32 * app * aqq:
        tau, a = aqq - app, 2.0 * apq
        t = (a if tau >= 0.0 else -a) / (abs(tau) + np.hypot(tau, a))
        c = 1.0 / np.sqrt(1.0 + t * t)
        return c, t * c
    return 1.0, 0.0


def givens(p, q, c, s):
    r = np.eye(8)
    r[p, p] = c; r[q, q] = c; r[p, q] = s; r[q, p] = -s
    return r


def dmma(c0, c1, a, b):
    """mma.sync.m8n8k4.f64: lane (g = lane >> 2, t = lane & 3) supplies A[g][t] and B[t][g], owns D[g][2t], D[g][2t + 1]."""
    A = np.zeros((8, 4)); B = np.zeros((4, 8))
    for l in LANES:
        A[l >> 2, l & 3] = a[l]; B[l & 3, l >> 2] = b[l]
    D = A @ B
    for l in LANES:
        c0[l] += D[l >> 2, 2 * (l & 3)]; c1[l] += D[l >> 2, 2 * (l & 3) + 1]


def warp_simultaneous(M):
    """The kernel's sequence for one block pair, lane by lane.  Returns J and J^T M J as the warp leaves them in shared memory."""
    Msh = np.zeros(8 * KJS); Jsh = np.zeros(8 * KJS)
    for i in range(8):
        for j in range(8):
            Msh[i * KJS + j] = M[i, j]
    slot = [l & 3 for l in LANES]; idx = [l >> 2 for l in LANES]; l4 = idx; l3 = slot
    # one rotation per lane: pair (slot, 4 + (slot + idx) % 4)
    c1 = np.zeros(32); s1 = np.zeros(32)
    for l in LANES:
        q = 4 + ((slot[l] + idx[l]) & 3)
        c1[l], s1[l] = rotation(Msh[slot[l] * (KJS + 1)], Msh[q * (KJS + 1)], Msh[slot[l] * KJS + q])
    # (c, s) of round r of a slot come from the lane 4 r + slot
    cr = np.array([[c1[4 * r + slot[l]] for r in range(4)] for l in LANES])
    sr = np.array([[s1[4 * r + slot[l]] for r in range(4)] for l in LANES])
    jp = np.array([1.0 if idx[l] == slot[l] else 0.0 for l in LANES])
    jq = np.array([1.0 if idx[l] == 4 + slot[l] else 0.0 for l in LANES])
    src = [(l & ~3) | ((l + 1) & 3) for l in LANES]
    for r in range(4):
        np_, nq = cr[:, r] * jp - sr[:, r] * jq, sr[:, r] * jp + cr[:, r] * jq
        jp = np_
        jq = nq[src] if r < 3 else nq
    for l in LANES:
        Jsh[idx[l] * KJS + slot[l]] = jp[l]
        Jsh[idx[l] * KJS + 4 + ((slot[l] + 3) & 3)] = jq[l]
    # X = M J, then M' = J^T X
    x0 = np.zeros(32); x1 = np.zeros(32)
    dmma(x0, x1, [Msh[l4[l] * KJS + l3[l]] for l in LANES], [Jsh[l3[l] * KJS + l4[l]] for l in LANES])
    dmma(x0, x1, [Msh[l4[l] * KJS + 4 + l3[l]] for l in LANES], [Jsh[(4 + l3[l]) * KJS + l4[l]] for l in LANES])
    for l in LANES:
        Msh[l4[l] * KJS + 2 * l3[l]] = x0[l]; Msh[l4[l] * KJS + 2 * l3[l] + 1] = x1[l]
    y0 = np.zeros(32); y1 = np.zeros(32)
    dmma(y0, y1, [Jsh[l3[l] * KJS + l4[l]] for l in LANES], [Msh[l3[l] * KJS + l4[l]] for l in LANES])
    dmma(y0, y1, [Jsh[(4 + l3[l]) * KJS + l4[l]] for l in LANES], [Msh[(4 + l3[l]) * KJS + l4[l]] for l in LANES])
    for l in LANES:
        Msh[l4[l] * KJS + 2 * l3[l]] = y0[l]; Msh[l4[l] * KJS + 2 * l3[l] + 1] = y1[l]
    J = np.array([[Jsh[i * KJS + j] for j in range(8)] for i in range(8)])
    Mn = np.array([[Msh[i * KJS + j] for j in range(8)] for i in range(8)])
    return J, Mn


def test_lane_mapping_of_the_simultaneous_inner_solve():
    rng = np.random.default_rng(1)
    nsmall = 0
    for trial in range(5):
        d = np.sort(rng.uniform(0.5, 1.5, 8))
        E = 10.0 ** -(2 + trial) * rng.standard_normal((8, 8))
        M = np.diag(d) + E + E.T
        J, Mn = warp_simultaneous(M)
        # J = product over the rounds r of the four disjoint rotations (slot, 4 + (slot + r) % 4), all angles from M itself
        Jref = np.eye(8)
        for r in range(4):
            for s in range(4):
                q = 4 + ((s + r) & 3)
                Jref = Jref @ givens(s, q, *rotation(M[s, s], M[q, q], M[s, q]))
        assert np.max(np.abs(J - Jref)) < 1e-15
        assert np.max(np.abs(J.T @ J - np.eye(8))) < 2e-15
        assert np.max(np.abs(Mn - J.T @ M @ J)) < 1e-15
        # the cross block shrinks like that of the sequential rounds (angles re-evaluated after every round)
        Ms = M.copy()
        for r in range(4):
            Rr = np.eye(8)
            for s in range(4):
                q = 4 + ((s + r) & 3)
                Rr = Rr @ givens(s, q, *rotation(Ms[s, s], Ms[q, q], Ms[s, q]))
            Ms = Rr.T @ Ms @ Rr
        before, sim, seq = (np.max(np.abs(X[:4, 4:])) for X in (M, Mn, Ms))
        max_sin = max(abs(rotation(M[s, s], M[q, q], M[s, q])[1]) for s in range(4) for q in range(4, 8))
        if max_sin <= 0.1:      # the kernel's condition for the simultaneous form
            assert sim < 0.01 * before and sim < 3.0 * seq + 1e-15, (trial, before, sim, seq)
            nsmall += 1
    assert nsmall >= 3


def _block_pair(step, pr, nb):
    """jb_block_pair as the kernel evaluates it: residues modulo nb - 1 by one conditional subtraction / addition."""
    mm1 = nb - 1
    if pr == 0:
        p, q = step, mm1
    else:
        p = step + pr
        if p >= mm1:
            p -= mm1
        q = step - pr
        if q < 0:
            q += mm1
    return min(p, q), max(p, q)


def test_block_pairs_are_a_round_robin_tournament():
    """Every sweep pairs every two column blocks exactly once, every step uses every block exactly once, and the division-free
    indices are those of the circle method ((step + pr) mod (nb - 1), (step - pr) mod (nb - 1), block nb - 1 fixed)."""
    for nb in range(2, 42, 2):
        seen = set()
        for step in range(nb - 1):
            used = []
            for pr in range(nb // 2):
                P, Q = _block_pair(step, pr, nb)
                mm1 = nb - 1
                ref = (step % mm1, mm1) if pr == 0 else ((step + pr) % mm1, (step - pr + mm1) % mm1)
                assert (P, Q) == (min(ref), max(ref))
                assert P < Q and (P, Q) not in seen
                seen.add((P, Q))
                used += [P, Q]
            assert sorted(used) == list(range(nb))
        assert len(seen) == nb * (nb - 1) // 2


_PAIRTAB = [[(s7, 7) if slot == 0 else tuple(sorted(((s7 + slot) % 7, (s7 - slot + 7) % 7))) for slot in range(4)]
            for s7 in range(7)]


def block_jacobi(G, tol, qtol, max_sin, maxsweeps=40):
    """jb_sweeps2 at the level of matrices: blocks of four columns, round-robin tournament, full 8 x 8 Gram matrix and
    tournament at a sweep's first step, tracked block Gram matrices and P x Q rotations otherwise, quadratic stop."""
    n = G.shape[1]
    nb = n // 4
    T = {}
    stats = dict(simultaneous=0, sequential=0)
    sweep = 0
    while sweep < maxsweeps:
        rotated = rotbig = False
        for step in range(nb - 1):
            full = step == 0
            for pr in range(nb // 2):
                P, Q = _block_pair(step, pr, nb)
                cols = list(range(4 * P, 4 * P + 4)) + list(range(4 * Q, 4 * Q + 4))
                W = G[:, cols]
                if full:
                    M = W.T @ W
                else:
                    C = W[:, :4].T @ W[:, 4:]
                    M = np.block([[T[P], C], [C.T, T[Q]]])
                dg = np.diag(M)
                R = M * M / np.outer(dg, dg)
                np.fill_diagonal(R, 0.0)
                if not (R > tol * tol).any():
                    if full:
                        T[P], T[Q] = M[:4, :4].copy(), M[4:, 4:].copy()
                    continue
                rotated = True
                rotbig = rotbig or bool((R > qtol * qtol).any())
                J = np.eye(8)
                cs = {(s, r): rotation(M[s, s], M[4 + ((s + r) & 3), 4 + ((s + r) & 3)], M[s, 4 + ((s + r) & 3)])
                      for s in range(4) for r in range(4)}
                if not full and max_sin > 0.0 and all(abs(v[1]) <= max_sin for v in cs.values()):
                    for r in range(4):
                        for s in range(4):
                            J = J @ givens(s, 4 + ((s + r) & 3), *cs[(s, r)])
                    M = J.T @ M @ J
                    stats["simultaneous"] += 1
                else:
                    rounds = _PAIRTAB if full else [[(s, 4 + ((s + r) & 3)) for s in range(4)] for r in range(4)]
                    for row in rounds:
                        Rr = np.eye(8)
                        for (p, q) in row:
                            Rr = Rr @ givens(p, q, *rotation(M[p, p], M[q, q], M[p, q]))
                        M = Rr.T @ M @ Rr
                        J = J @ Rr
                    stats["sequential"] += 1
                T[P], T[Q] = M[:4, :4].copy(), M[4:, 4:].copy()
                G[:, cols] = W @ J
        sweep += 1
        if not rotated or not rotbig:
            break
    return sweep, stats


def _warm_problem(n, pert, seed):
    rng = np.random.default_rng(seed)
    lam = np.sort(np.concatenate([rng.uniform(-25.0, -8.0, n // 2), rng.uniform(-2.0, 30.0, n - n // 2)]))
    Qm, _ = np.linalg.qr(rng.standard_normal((n, n)))
    A = Qm @ np.diag(lam) @ Qm.T
    A = 0.5 * (A + A.T)
    Z0 = np.linalg.eigh(A)[1]
    D = rng.standard_normal((n, n))
    A2 = A + 0.5 * (D + D.T) * pert
    sigma = 1.05 * np.abs(A2).sum(1).max() + 1.0
    return A2, sigma, (A2 @ Z0 + sigma * Z0) / sigma


def test_simultaneous_rotations_keep_the_sweep_counts():
    """Warm-started problems as the SCC iterations pose them (G = (A' + sigma) Z_prev / sigma with A' perturbed by 3e-2 ... 3e-5):
    with |sin| <= 0.1 the simultaneous form takes over nine pairs out of ten, the sweep count and the spectrum stay."""
    n = 48
    tol = 8.0 * np.sqrt(n) * 2.220446049250313e-16
    for pert in (3e-2, 3e-3, 3e-5):
        for seed in range(2):
            A2, sigma, G0 = _warm_problem(n, pert, seed)
            w = np.linalg.eigvalsh(A2)
            res = {}
            for max_sin in (0.0, 0.1):
                G = G0.copy()
                sweeps, stats = block_jacobi(G, tol, 1e-7, max_sin)
                nrm = np.linalg.norm(G, axis=0)
                Z = G / nrm
                assert np.max(np.abs(np.sort(sigma * (nrm - 1.0)) - w)) < 5e-12 * sigma
                assert np.max(np.abs(Z.T @ Z - np.eye(n))) < 1e-13
                res[max_sin] = (sweeps, stats)
            assert res[0.1][0] == res[0.0][0], (pert, seed, res)
            st = res[0.1][1]
            assert st["simultaneous"] > 5 * st["sequential"], (pert, seed, st)
