"""Fixture for the solver's distance tables: runs the REFERENCE's own
paikin_tal_solver.inter_piece_distance.InterPieceDistance (imports unchanged from /root/reference) with the distance
callback of evaluation.py:116-131 on seeded [N, N, 4] distance arrays, and stores every table its constructor fills.
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


def reference_tables(d, order):
    sys.path.insert(0, REF)
    from paikin_tal_solver.inter_piece_distance import InterPieceDistance
    from paikin_tal_solver.puzzle_piece import PuzzlePieceSide
    from paikin_tal_solver.puzzle_importer import PuzzleType

    def distance_function(piece_i, piece_i_side, piece_j, piece_j_side):   # evaluation.py:116-131, verbatim semantics
        pred = d[piece_i.origin_piece_id][piece_j.origin_piece_id]
        if piece_j_side == PuzzlePieceSide.left:
            if piece_i_side == PuzzlePieceSide.right:
                return pred[0] * 1000.
        if piece_j_side == PuzzlePieceSide.right:
            if piece_i_side == PuzzlePieceSide.left:
                return pred[2] * 1000.
        if piece_j_side == PuzzlePieceSide.top:
            if piece_i_side == PuzzlePieceSide.bottom:
                return pred[1] * 1000.
        if piece_j_side == PuzzlePieceSide.bottom:
            if piece_i_side == PuzzlePieceSide.top:
                return pred[3] * 1000.
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
        for k, v in ref.items():
            blob[f'{k}_{seed}'] = v
        print(case, 'ties:', int((ref['candidates'].sum(-1) > 1).sum()), 'best buddies:', int((ref['best_buddy'] >= 0).sum()),
              'zeros:', int((ref['asym_dist'] == 0).sum()), 'start[0..2]:', ref['start_order'][:3].tolist())
    np.savez_compressed(os.path.join(HERE, 'solver_tables.npz'), **blob)
