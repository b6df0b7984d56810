"""Fixture for the solver's distance tables: runs the REFERENCE's own
paikin_tal_solver.inter_piece_distance.InterPieceDistance (imports unchanged from /root/reference) with the distance
callback of evaluation.py:116-131 on seeded [N, N, 4] distance arrays, and stores every table its constructor fills.

Scalar rules: the closure returns ``pred[k] * 1000.`` (np.float32 scalar x Python float). The reference's pinned stack
(requirements.txt: torch~=2.1, scipy~=1.9.1, scikit-image~=0.20) runs on NumPy 1.x, where that product is float64;
this container has NumPy 2.3 (NEP 50), where it stays float32. Both fixture sets are written by the reference's own
class: keys ``<table>_<seed>`` with the closure evaluated natively here (NumPy >= 2 rules) and ``<table>_np1_<seed>``
with the product spelled ``np.float64(pred[k]) * 1000.`` -- the value NumPy 1.x's promotion computes. The two sets
differ in the quantised cases (seeds 4 and 6: hundreds of distances off by one, e.g. float32(0.02) * 1000 truncates to
19 in float64 and to 20 in float32) and in a handful of entries of the continuous ones.
Run in the build container: python tests/golden/make_golden_tables.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'

# (seed, grid rows, grid cols, quantum x 1000, zero rate x 1000): distances are multiples of `quantum` (0 = continuous)
# so that ties occur; a fraction is forced to exactly 0 (the reference special-cases dist == 0 and second best == 0)
CASES = [(0, 3, 4, 0, 0), (1, 3, 3, 4, 30), (2, 1, 2, 0, 0), (3, 5, 6, 0, 2), (4, 4, 4, 20, 30), (5, 1, 3, 250, 0),
         (6, 6, 7, 1, 1)]


def case_inputs(seed, rows, cols, quantum_milli, zero_milli):
    """distance [n, n, 4] fp32 in [0, 1) indexed by origin piece id (what ``1 - sigmoid(logits)`` looks like: true
    neighbours of a rows x cols puzzle get small distances in the matching bin, evaluation.py:118-129 bin order), and
    the shuffled piece order (position -> origin id) that random.shuffle gives at evaluation.py:87."""
    rng = np.random.default_rng(seed)
    n = rows * cols
    d = (0.2 + 0.8 * rng.random((n, n, 4))).astype(np.float32)
    for r in range(rows):
        for c in range(cols):
            i = r * cols + c
            if c + 1 < cols:                                   # j is i's right neighbour
                d[i, i + 1, 0] = 0.15 * rng.random()           # i.right - j.left
                d[i + 1, i, 2] = 0.15 * rng.random()           # j.left - i.right
            if r + 1 < rows:                                   # j is i's bottom neighbour
                d[i, i + cols, 1] = 0.15 * rng.random()
                d[i + cols, i, 3] = 0.15 * rng.random()
    if quantum_milli > 0:
        q = quantum_milli / 1000.0
        d = (np.floor(d / q) * q).astype(np.float32)
    if zero_milli > 0:
        d[rng.random((n, n, 4)) < zero_milli / 1000.0] = 0.0
    order = rng.permutation(n).astype(np.int32)
    return d, order


class _Piece:
    def __init__(self, origin):
        self.origin_piece_id = int(origin)
        self.id_number = -1


def reference_tables(d, order, numpy1=False):
    sys.path.insert(0, REF)
    times1000 = (lambda v: np.float64(v) * 1000.) if numpy1 else (lambda v: v * 1000.)
    from paikin_tal_solver.inter_piece_distance import InterPieceDistance
    from paikin_tal_solver.puzzle_piece import PuzzlePieceSide
    from paikin_tal_solver.puzzle_importer import PuzzleType

    def distance_function(piece_i, piece_i_side, piece_j, piece_j_side):   # evaluation.py:116-131, verbatim semantics
        pred = d[piece_i.origin_piece_id][piece_j.origin_piece_id]
        if piece_j_side == PuzzlePieceSide.left:
            if piece_i_side == PuzzlePieceSide.right:
                return times1000(pred[0])
        if piece_j_side == PuzzlePieceSide.right:
            if piece_i_side == PuzzlePieceSide.left:
                return times1000(pred[2])
        if piece_j_side == PuzzlePieceSide.top:
            if piece_i_side == PuzzlePieceSide.bottom:
                return times1000(pred[1])
        if piece_j_side == PuzzlePieceSide.bottom:
            if piece_i_side == PuzzlePieceSide.top:
                return times1000(pred[3])
        return float('inf')

    pieces = [_Piece(o) for o in order]
    ipd = InterPieceDistance(pieces, distance_function, PuzzleType.type1)
    n = len(order)
    info = ipd._piece_distance_info
    out = {
        'asym_dist': np.stack([p._asymmetric_distances[:, :, 0] for p in info]),            # [n, 4, n] uint32
        'asym_compat': np.stack([p._asymmetric_compatibilities[:, :, 0] for p in info]),    # [n, 4, n] float32
        'mutual_compat': np.stack([p._mutual_compatibilities[:, :, 0] for p in info]),      # [n, 4, n] float32
        'min_dist': np.array([[int(v) for v in p._min_distance] for p in info], dtype=np.int64),
        'second_dist': np.array([[int(v) for v in p._second_best_distance] for p in info], dtype=np.int64),
    }
    cand = np.zeros((n, 4, n), dtype=bool)
    bb = np.full((n, 4), -1, dtype=np.int32)
    for i, p in enumerate(info):
        for s in range(4):
            js = [j for (j, side) in p._best_buddy_candidates[s]]
            assert js == sorted(js) and all(side.value == (s + 2) % 4 for (_, side) in p._best_buddy_candidates[s])
            cand[i, s, js] = True
            assert len(p._best_buddies[s]) <= 1
            if p._best_buddies[s]:
                bb[i, s] = p._best_buddies[s][0][0]
    out['candidates'] = cand
    out['best_buddy'] = bb
    out['start_order'] = np.array([(a, b) for (a, b, _) in ipd._start_piece_ordering], dtype=np.int64).reshape(-1, 2)
    out['start_compat'] = np.array([float(c) for (_, _, c) in ipd._start_piece_ordering], dtype=np.float64)
    return out


if __name__ == '__main__':
    blob = {'cases': np.array(CASES, dtype=np.int64)}
    for case in CASES:
        seed = case[0]
        d, order = case_inputs(*case)
        ref = reference_tables(d, order)
        ref1 = reference_tables(d, order, numpy1=True)
        for k, v in ref.items():
            blob[f'{k}_{seed}'] = v
        for k, v in ref1.items():
            blob[f'{k}_np1_{seed}'] = v
        print('  entries where the NumPy 1.x / >= 2 distances differ:', int((ref['asym_dist'] != ref1['asym_dist']).sum()))
        print(case, 'ties:', int((ref['candidates'].sum(-1) > 1).sum()), 'best buddies:', int((ref['best_buddy'] >= 0).sum()),
              'zeros:', int((ref['asym_dist'] == 0).sum()), 'start[0..2]:', ref['start_order'][:3].tolist())
    np.savez_compressed(os.path.join(HERE, 'solver_tables.npz'), **blob)
