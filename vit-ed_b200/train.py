"""Host side of the Hisfrag training step. Only the pair construction exists so far: the backward
kernels that a training step needs are not built, and ``train_step`` says so instead of falling back to PyTorch."""
import torch

from . import _lib


def prepare_pairs(targets, generator=None):
    """(groups [P, 2] int64, labels [P, 1] fp32) of one batch, as HisfragTrainer.prepare_data builds them
    (hisfrag.py:117-147): for every item i the later items j > i of the same writer are positive pairs (i, j), those of
    another writer negative pairs; at most twice as many negatives as positives survive a ``torch.randperm`` draw
    (``generator`` seeds it; None = the global RNG as in the reference); positives first. The decoder then scores
    ``model(tokens[groups[:, 1]], images[groups[:, 0]])`` (:152-159)."""
    t = torch.as_tensor(targets).view(-1).cpu()
    n = t.numel()
    same = t.view(-1, 1) == t.view(1, -1)
    upper = torch.triu(torch.ones(n, n, dtype=torch.bool), diagonal=0)
    pos = torch.nonzero(same & torch.triu(torch.ones(n, n, dtype=torch.bool), diagonal=1))   # row-major = i-major, j ascending
    neg = torch.nonzero(~same & upper)
    keep = min(neg.shape[0], int(2 * pos.shape[0]))
    perm = torch.randperm(neg.shape[0], generator=generator) if generator is not None else torch.randperm(neg.shape[0])
    neg = neg[perm[:keep]]
    labels = torch.cat([torch.ones(pos.shape[0]), torch.zeros(neg.shape[0])]).view(-1, 1)
    return torch.cat([pos, neg], dim=0), labels


def train_step(*args, **kwargs):
    raise _lib.VitedError('the Hisfrag training step (SURVEY 8f row 1) is not built: it needs backward kernels; '
                          'oracle.train_step and tests/golden/train_step.npz hold its reference results')
