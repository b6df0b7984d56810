"""Distance tables of the Paikin-Tal solver straight from the score matrix.

The reference hands the solver a Python closure (evaluation.py:116-131); ``InterPieceDistance.__init__``
(paikin_tal_solver/inter_piece_distance.py:437-475) then calls it 4*N*(N-1) times and runs three more O(N^2) Python
loops before the first piece is placed. ``build_tables`` computes the same tables -- bit-identical -- on the device
from the [N, N, 4] logits that ``grid.score_puzzle`` returns (C-ABI ``vited_puzzle_tables``); ``install`` /
``factory`` wrap them in the reference's own ``InterPieceDistance`` object so the solver runs unmodified.
"""
import ctypes
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib

MAXSIZE = 2 ** 63 - 1   # sys.maxsize: the reference's initial second-best distance (inter_piece_distance.py:283-288)


@dataclass
class PuzzleTables:
    """Everything ``InterPieceDistance.__init__`` computes, indexed by LIST POSITION of the piece (the reference numbers
    pieces by position, inter_piece_distance.py:437-441) and by PuzzlePieceSide value (top 0, right 1, bottom 2, left 3)."""
    n: int
    asym_dist: np.ndarray        # [N, 4, N] uint32, diagonal 2^31 - 1
    min_dist: np.ndarray         # [N, 4] int64
    second_dist: np.ndarray      # [N, 4] int64
    n_candidates: np.ndarray     # [N, 4] int32: pieces at the minimum distance
    candidate: np.ndarray        # [N, 4] int32: the lowest such piece, -1 if none
    asym_compat: np.ndarray      # [N, 4, N] float32, diagonal +inf
    mutual_compat: np.ndarray    # [N, 4, N] float32, diagonal +inf
    best_buddy: np.ndarray       # [N, 4] int32: piece or -1
    start_piece_ordering: list = field(default_factory=list)   # [(piece, numb_bb_neighbors, total_compatibility)]


def build_tables(scores, order=None, scores_are_logits=True, scalar_rules='numpy1'):
    """scores: CUDA fp32 tensor [N, N, 4] indexed by origin piece id -- the logits of ``grid.score_puzzle``
    (``scores_are_logits=True``: 1 - sigmoid is applied on the device, evaluation.py:109-114) or distances.
    order[k]: origin id of the piece at list position k (evaluation.py:87 shuffles), None = identity.
    scalar_rules: how the closure's ``pred[k] * 1000.`` (np.float32 scalar x Python float) is evaluated --
    'numpy1' (default): float64, as under the NumPy 1.x the reference's pinned requirements run on; 'numpy2': float32
    (NEP 50), what the same reference code computes when run under NumPy >= 2. The truncated uint32 distances differ
    by one on ~1e-5 of the entries."""
    if scalar_rules not in ('numpy1', 'numpy2'):
        raise _lib.VitedError(f"build_tables: scalar_rules must be 'numpy1' or 'numpy2', got {scalar_rules!r}")
    flags = (1 if scores_are_logits else 0) | (2 if scalar_rules == 'numpy2' else 0)
    if not (isinstance(scores, torch.Tensor) and scores.is_cuda):
        raise _lib.VitedError('build_tables: scores must be a CUDA tensor (there is no CPU path)')
    if scores.dtype != torch.float32 or scores.dim() != 3 or scores.shape[0] != scores.shape[1] or scores.shape[2] != 4:
        raise _lib.VitedError(f'build_tables: expected fp32 [N, N, 4], got {scores.dtype} {tuple(scores.shape)}')
    scores = scores.contiguous()
    n = scores.shape[0]
    dev = scores.device
    order_t = None
    if order is not None:
        order_np = np.asarray(order, dtype=np.int32)
        if sorted(order_np.tolist()) != list(range(n)):
            raise _lib.VitedError('build_tables: order must be a permutation of range(N)')
        order_t = torch.from_numpy(order_np).to(dev)
    with torch.cuda.device(dev):
        asym = torch.empty((n, 4, n), dtype=torch.int32, device=dev)       # uint32 bit patterns
        compat = torch.empty((n, 4, n), dtype=torch.float32, device=dev)
        mutual = torch.empty((n, 4, n), dtype=torch.float32, device=dev)
        min_d = torch.empty((n, 4), dtype=torch.int64, device=dev)
        second_d = torch.empty((n, 4), dtype=torch.int64, device=dev)
        n_cand = torch.empty((n, 4), dtype=torch.int32, device=dev)
        cand = torch.empty((n, 4), dtype=torch.int32, device=dev)
        bb = torch.empty((n, 4), dtype=torch.int32, device=dev)
        p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.lib.vited_puzzle_tables(p(scores), flags, p(order_t), n, p(asym), p(min_d),
                                                p(second_d), p(n_cand), p(cand), p(compat), p(mutual), p(bb), stream),
                   'vited_puzzle_tables')
        tables = PuzzleTables(
            n=n, asym_dist=asym.cpu().numpy().view(np.uint32), min_dist=min_d.cpu().numpy(),
            second_dist=second_d.cpu().numpy(), n_candidates=n_cand.cpu().numpy(), candidate=cand.cpu().numpy(),
            asym_compat=compat.cpu().numpy(), mutual_compat=mutual.cpu().numpy(), best_buddy=bb.cpu().numpy())
    tables.start_piece_ordering = start_piece_ordering(tables.best_buddy, tables.mutual_compat)
    return tables


def start_piece_ordering(best_buddy, mutual_compat):
    """InterPieceDistance.find_start_piece_candidates (inter_piece_distance.py:650-719): per piece 4 x its best buddies
    + the best buddies of those, then the summed mutual compatibility (fp32, side order); descending, stable."""
    n = best_buddy.shape[0]
    info = []
    for i in range(n):
        ids, total = [], 0
        for s in range(4):
            j = int(best_buddy[i, s])
            if j >= 0:
                ids.append(j)
                total = total + mutual_compat[i, s, j]
        info.append((ids, total))
    ordering = [(i, 4 * len(info[i][0]) + sum(len(info[b][0]) for b in info[i][0]), info[i][1]) for i in range(n)]
    ordering.sort(key=lambda t: (t[1], t[2]), reverse=True)
    return ordering


def install(tables, pieces, puzzle_type, inter_piece_distance_cls, piece_info_cls, side_enum):
    """An ``InterPieceDistance`` instance in the state its constructor leaves it in, without the callbacks. The three
    classes are the reference's own (paikin_tal_solver.inter_piece_distance.InterPieceDistance /
    PieceDistanceInformation, paikin_tal_solver.puzzle_piece.PuzzlePieceSide), passed in so that this package does not
    import the reference."""
    n = tables.n
    if len(pieces) != n:
        raise _lib.VitedError(f'install: {len(pieces)} pieces for tables of {n}')
    ipd = inter_piece_distance_cls.__new__(inter_piece_distance_cls)
    _fill(ipd, tables, pieces, puzzle_type, piece_info_cls, side_enum)
    return ipd


def _fill(ipd, tables, pieces, puzzle_type, piece_info_cls, side_enum):
    n = tables.n
    sides = side_enum.get_all_sides()
    for k, piece in enumerate(pieces):                       # inter_piece_distance.py:437-441
        piece.id_number = k
    ipd._numb_pieces = n
    ipd._distance_function = None                            # cleared at the end of the constructor (:474-475)
    ipd._puzzle_type = puzzle_type
    ipd._piece_distance_info = []
    for i in range(n):
        info = piece_info_cls(i, n, puzzle_type)
        info._asymmetric_distances = tables.asym_dist[i][:, :, None]          # (4, N, 1) as for a type-1 puzzle
        info._asymmetric_compatibilities = tables.asym_compat[i][:, :, None]
        info._mutual_compatibilities = tables.mutual_compat[i][:, :, None]
        as_ref = lambda v: int(v) if v >= 2 ** 32 else np.uint32(v)            # untouched initial values stay ints
        info._min_distance = [as_ref(v) for v in tables.min_dist[i]]
        info._second_best_distance = [as_ref(v) for v in tables.second_dist[i]]
        for s in sides:
            cs = s.complementary_side
            k = int(tables.n_candidates[i, s.value])
            if k == 1:
                js = [int(tables.candidate[i, s.value])]
            elif k > 1:
                row = tables.asym_dist[i, s.value]
                js = [int(j) for j in np.nonzero(row == np.uint32(tables.min_dist[i, s.value]))[0] if j != i]
            else:
                js = []
            info._best_buddy_candidates[s.value] = [(j, cs) for j in js]
            b = int(tables.best_buddy[i, s.value])
            info._best_buddies[s.value] = [(b, cs)] if b >= 0 else []
        ipd._piece_distance_info.append(info)
    ipd._start_piece_ordering = list(tables.start_piece_ordering)


def factory(tables, inter_piece_distance_cls, piece_info_cls, side_enum):
    """A stand-in for the ``InterPieceDistance`` name in paikin_tal_solver/solver.py:212: a subclass (the solver also
    calls the class's static helpers through that name) whose constructor takes the same arguments and fills the
    object from the prebuilt tables instead of calling ``distance_function``."""
    class PrebuiltInterPieceDistance(inter_piece_distance_cls):
        def __init__(self, pieces, distance_function, puzzle_type):
            if len(pieces) != tables.n:
                raise _lib.VitedError(f'{len(pieces)} pieces for tables of {tables.n}')
            _fill(self, tables, pieces, puzzle_type, piece_info_cls, side_enum)
    return PrebuiltInterPieceDistance
