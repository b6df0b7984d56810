"""Weights whose Hisfrag similarity logits SEPARATE writers, for the north-star retrieval clause ("identical retrieval
top-1 / mAP to 3 decimals"): random-init logits differ less between fragment pairs than any 16-bit noise, and no
trained checkpoint exists offline. A production-shaped model (embed_dim 384, 6 heads of 64, 256 patch tokens: the
fused Linear+residual+LayerNorm kernels and the tcgen05 long-sequence attention run on it) gets seeded synthetic
weights; only its vectors -- biases, LayerNorm weights, class token, head -- are then trained with the reference's own
training objective (hisfrag.py:117-159: positive / negative pairs of a batch, BCE-with-logits), by autograd over the
oracle's functional model on CPU, on synthetic fragments with a writer signal. The fixture stores the trained vectors
only (the matrices are regenerated from the seed), so it stays small.

  python tests/golden/make_retrieval_fixture.py        (a few minutes of CPU)
"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

KW = dict(img_size=256, patch_size=16, num_classes=1, embed_dim=384, depth=1, c_depth=2, num_heads=6)
WEIGHT_SEED, WRITER_SEED, N_WRITERS = 7, 11, 8
STEPS, LR = int(os.environ.get('STEPS', 120)), 4e-3


def base_state_dict():
    from vited_b200 import synthetic   # (pure-Python helper; nothing here touches the CUDA library's compute)
    shapes = synthetic.state_dict_shapes(**KW)
    return synthetic.synthetic_state_dict(shapes, seed=WEIGHT_SEED)


def trainable(key, t):
    return t.dim() == 1 or key in ('cls_token', 'head.weight')


MIXED = dict(n_writers=12, per_writer=4, seed=6)   # 8 writers seen in training + 4 unseen ones: metrics below 1


def mixed_set():
    from vited_b200 import synthetic
    return synthetic.synthetic_fragments(MIXED['n_writers'], MIXED['per_writer'], KW['img_size'], seed=MIXED['seed'],
                                         writer_seed=WRITER_SEED)


def add_mixed_only():
    """Re-evaluates the stored vectors on the mixed set without retraining (python make_retrieval_fixture.py mixed)."""
    from oracle import vited_oracle as orc
    path = os.path.join(HERE, 'retrieval_weights.npz')
    z = dict(np.load(path))
    sd = base_state_dict()
    for k, v in z.items():
        if not k.startswith('__'):
            sd[k] = torch.from_numpy(v)
    imgs, labels = mixed_set()
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        sim = orc.score_fragment_grid(sd, KW['num_heads'], imgs)
    m = orc.wi19_metrics(orc.sim_to_distance(sim), labels.numpy(), kind='stable')
    print('mixed set (8 seen + 4 unseen writers x 4): mAP %.4f top-1 %.4f Pr@10 %.4f Pr@100 %.4f' % m)
    z['__mixed_sim__'] = sim.numpy()
    z['__mixed_metrics__'] = np.array(m, dtype=np.float64)
    np.savez_compressed(path, **z)


def main():
    from oracle import vited_oracle as orc
    from vited_b200 import synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    sd = base_state_dict()
    params = {k: (v.clone().requires_grad_(True) if trainable(k, v) else v) for k, v in sd.items()}
    train = [p for k, p in params.items() if p.requires_grad]
    opt = torch.optim.Adam(train, lr=LR)
    H = KW['num_heads']
    t0 = time.time()
    for step in range(STEPS):
        imgs, labels = synthetic.synthetic_fragments(N_WRITERS, 2, KW['img_size'], seed=1000 + step, writer_seed=WRITER_SEED)
        # shuffled batch, as a DataLoader delivers it: with the writer-major order the pairs (i, j > i) would always put
        # the higher-numbered writer on the context side and the vectors would learn that order instead of similarity
        perm = torch.randperm(len(labels), generator=torch.Generator().manual_seed(step))
        imgs, labels = imgs[perm], labels[perm]
        groups, y = orc.train_pairs(labels, perm_seed=step)
        tokens = orc.forward_first_part(imgs, params, H)
        x2 = orc.forward_second_part(tokens[groups[:, 1]], imgs[groups[:, 0]], params, H)
        logits = orc.forward_head(x2, params)
        loss = F.binary_cross_entropy_with_logits(logits, y)
        opt.zero_grad()
        loss.backward()
        opt.step()
        if step % 10 == 0 or step == STEPS - 1:
            acc = ((logits > 0).float() == y).float().mean().item()
            print(f'step {step} loss {loss.item():.4f} acc {acc:.3f} ({time.time() - t0:.0f}s)', flush=True)
    final = {k: v.detach() for k, v in params.items()}
    imgs, labels = synthetic.synthetic_fragments(N_WRITERS, 6, KW['img_size'], seed=5, writer_seed=WRITER_SEED)
    with torch.no_grad():
        sim = orc.score_fragment_grid(final, H, imgs)
    m = orc.wi19_metrics(orc.sim_to_distance(sim), labels.numpy(), kind='stable')
    print('held-out 48 fragments: mAP %.4f top-1 %.4f Pr@10 %.4f Pr@100 %.4f' % m)
    same = labels[:, None] == labels[None, :]
    print('logits same-writer: min %.3f mean %.3f | other-writer: max %.3f mean %.3f' % (
        sim[same].min(), sim[same].mean(), sim[~same].max(), sim[~same].mean()))
    out = {k: v.numpy() for k, v in final.items() if trainable(k, v)}
    out['__kwargs__'] = np.array(repr(KW))
    out['__seeds__'] = np.array([WEIGHT_SEED, WRITER_SEED, N_WRITERS], dtype=np.int64)
    out['__heldout_metrics__'] = np.array(m, dtype=np.float64)
    out['__heldout_sim__'] = sim.numpy()
    np.savez_compressed(os.path.join(HERE, 'retrieval_weights.npz'), **out)
    print('wrote retrieval_weights.npz:', sum(v.size for k, v in out.items() if not k.startswith('__')), 'trained values')


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'mixed':
        add_mixed_only()
    else:
        main()
        add_mixed_only()
