"""ORACLE -- test infrastructure only. NOT part of the product path.

A plain fp32 CPU restatement of the reference algorithm for the all-pairs scoring path of glmanhtu/vit-ed. Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg may import it, and
only as the checker / the reported CPU baseline; nothing under ``vit-ed_b200/`` imports it.

Pinning (see tests/golden/make_golden.py and DESIGN.md): the reference publishes no golden vectors for this path
(SURVEY 8c). The oracle is pinned against outputs of the reference's OWN ``models/vision_transformer.py`` executed in
the build container (fixtures in tests/golden/*.npz) -- with one caveat: ``timm==0.9.2`` (requirements.txt:2) is
absent, so the five timm symbols that file imports were stood in for by tests/golden/_timm_shim.py, a restatement of
timm 0.9.2's published semantics. For the timm-owned arithmetic (PatchEmbed, Mlp, _pos_embed, forward_head,
LayerNorm eps) parity is therefore UNPINNED by any reference-held vector; for the reference-owned code (Attention,
Block, CrossAttention, CrossBlock, three-mode forward) and for the integer bookkeeping (pair enumeration, crop
geometry, sampler split: real reference code imported unchanged) it is pinned. tests/test_timm_semantics.py cross-checks
the timm-owned part (shim and oracle) against torchvision's VisionTransformer, an independent implementation of the
same forward, on the same weights.

Every function cites the reference file:line it restates.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-6  # VisionTransformerCustom: norm_layer = partial(nn.LayerNorm, eps=1e-6) (vision_transformer.py:348)
# The reference takes F.scaled_dot_product_attention when timm's use_fused_attn() says so (vision_transformer.py:32,
# :63-66); the explicit softmax below is its other branch (:68-72). Same arithmetic; bench.py's torch GPU baseline turns
# this on so that the stock fused kernels are what gets timed.
USE_SDPA = False


# ------------------------------------------------------------------------------------------------ model arithmetic
def _ln(x, sd, prefix):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + '.weight'], sd[prefix + '.bias'], EPS)


def _linear(x, sd, prefix):
    return F.linear(x, sd[prefix + '.weight'], sd.get(prefix + '.bias'))


def patch_embed(images, sd):
    """timm PatchEmbed: Conv2d(k=s=p) -> flatten(2).transpose(1,2) (called at vision_transformer.py:383,391).
    Written as im2col + matmul so the patch indexing contract is explicit."""
    w = sd['patch_embed.proj.weight']
    d, c, p, _ = w.shape
    b, _, s, _ = images.shape
    g = s // p
    cols = images.reshape(b, c, g, p, g, p).permute(0, 2, 4, 1, 3, 5).reshape(b, g * g, c * p * p)
    return cols @ w.reshape(d, c * p * p).t() + sd['patch_embed.proj.bias']


def attention(x, sd, prefix, num_heads):
    """Attention.forward (vision_transformer.py:56-80): qkv split (3, H, hd), softmax(q k^T * hd^-0.5) v, proj."""
    b, n, c = x.shape
    hd = c // num_heads
    qkv = _linear(x, sd, prefix + '.qkv').reshape(b, n, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    if USE_SDPA:
        y = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, n, c)
        return _linear(y, sd, prefix + '.proj')
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
    attn = attn.softmax(dim=-1)
    y = (attn @ v).transpose(1, 2).reshape(b, n, c)
    return _linear(y, sd, prefix + '.proj')


def cross_attention(x, context, sd, prefix, num_heads):
    """CrossAttention.forward (vision_transformer.py:174-200): q from x, kv split (2, H, hd) from context."""
    b, n, c = x.shape
    nc = context.shape[1]
    hd = c // num_heads
    q = _linear(x, sd, prefix + '.q').reshape(b, n, num_heads, hd).permute(0, 2, 1, 3)
    kv = _linear(context, sd, prefix + '.kv').reshape(b, nc, 2, num_heads, hd).permute(2, 0, 3, 1, 4)
    k, v = kv.unbind(0)
    if USE_SDPA:
        y = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, n, c)
        return _linear(y, sd, prefix + '.proj')
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
    attn = attn.softmax(dim=-1)
    y = (attn @ v).transpose(1, 2).reshape(b, n, c)
    return _linear(y, sd, prefix + '.proj')


def mlp(x, sd, prefix):
    """timm Mlp: fc1 -> GELU (exact erf) -> fc2 (vision_transformer.py:115-120, :259-264)."""
    return _linear(F.gelu(_linear(x, sd, prefix + '.fc1')), sd, prefix + '.fc2')


def block(x, sd, prefix, num_heads):
    """Block.forward (vision_transformer.py:124-127); LayerScale / DropPath are Identity."""
    x = x + attention(_ln(x, sd, prefix + '.norm1'), sd, prefix + '.attn', num_heads)
    x = x + mlp(_ln(x, sd, prefix + '.norm2'), sd, prefix + '.mlp')
    return x


def cross_block(x, context, sd, prefix, num_heads):
    """CrossBlock.forward (vision_transformer.py:268-272)."""
    x = x + attention(_ln(x, sd, prefix + '.norm1'), sd, prefix + '.attn', num_heads)
    x = x + cross_attention(_ln(x, sd, prefix + '.norm_cross'), _ln(context, sd, prefix + '.norm_context'), sd,
                            prefix + '.cross_attn', num_heads)
    x = x + mlp(_ln(x, sd, prefix + '.norm2'), sd, prefix + '.mlp')
    return x


def _depths(sd):
    depth = 1 + max(int(k.split('.')[1]) for k in sd if k.startswith('blocks.'))
    c_depth = 1 + max(int(k.split('.')[1]) for k in sd if k.startswith('cross_blocks.'))
    return depth, c_depth


def forward_first_part(images, sd, num_heads):
    """vision_transformer.py:382-388: patch_embed + pos_embed[:, 1:] (:378-380) + encoder blocks; no final norm."""
    depth, _ = _depths(sd)
    x = patch_embed(images, sd) + sd['pos_embed'][:, 1:]
    for l in range(depth):
        x = block(x, sd, f'blocks.{l}', num_heads)
    return x


def prepare_x2(images, sd):
    """vision_transformer.py:390-395: patch_embed, cls token prepended, + pos_embed (timm _pos_embed)."""
    x = patch_embed(images, sd)
    x = torch.cat([sd['cls_token'].expand(x.shape[0], -1, -1), x], dim=1)
    return x + sd['pos_embed']


def forward_second_part(x1, images2, sd, num_heads):
    """vision_transformer.py:397-405: cross blocks then final norm."""
    _, c_depth = _depths(sd)
    x2 = prepare_x2(images2, sd)
    for l in range(c_depth):
        x2 = cross_block(x2, x1, sd, f'cross_blocks.{l}', num_heads)
    return _ln(x2, sd, 'norm')


def forward_head(x, sd):
    """timm forward_head with global_pool='token': x[:, 0] -> fc_norm (Identity) -> head."""
    return _linear(x[:, 0], sd, 'head')


@torch.no_grad()
def forward(sd, num_heads, x, x2=None, first_part=False):
    """VisionTransformerCustom.forward (vision_transformer.py:412-420), three modes."""
    if first_part:
        return forward_first_part(x, sd, num_heads)
    if x2 is not None:
        return forward_head(forward_second_part(x, x2, sd, num_heads), sd)
    x1, xb = torch.unbind(x, 1)
    return forward_head(forward_second_part(forward_first_part(x1, sd, num_heads), xb, sd, num_heads), sd)


# ------------------------------------------------------------------------------------------------ grid drivers
@torch.no_grad()
def score_puzzle_grid(sd, num_heads, images, batch=64):
    """evaluation.py:101-114 restated with the two-phase path: logits[i, j] for every ordered pair i != j."""
    n = images.shape[0]
    tokens = torch.cat([forward(sd, num_heads, images[i:i + batch], first_part=True) for i in range(0, n, batch)])
    pairs = ordered_pairs(n)
    out = torch.zeros((n, n, sd['head.weight'].shape[0]), dtype=torch.float32)
    for p0 in range(0, len(pairs), batch):
        sub = pairs[p0:p0 + batch]
        logits = forward(sd, num_heads, tokens[sub[:, 0]], images[sub[:, 1]])
        out[sub[:, 0], sub[:, 1]] = logits
    return out


@torch.no_grad()
def score_fragment_grid(sd, num_heads, images, batch=16):
    """hisfrag.py:189-231 + :281-292 restated: symmetric raw-logit similarity matrix from the a<=b pair set."""
    n = images.shape[0]
    tokens = torch.cat([forward(sd, num_heads, images[i:i + batch], first_part=True) for i in range(0, n, batch)])
    pairs = upper_tri_pairs(n)
    sim = torch.zeros((n, n), dtype=torch.float32)
    for p0 in range(0, len(pairs), batch):
        sub = pairs[p0:p0 + batch]
        logits = forward(sd, num_heads, tokens[sub[:, 0]], images[sub[:, 1]])[:, 0]
        sim[sub[:, 0], sub[:, 1]] = logits
        sim[sub[:, 1], sub[:, 0]] = logits
    return sim


# ------------------------------------------------------------------------------------------------ integer bookkeeping
def ordered_pairs(n):
    """data/datasets/pieces_dataset.py:27-32 (nested loops, skip i == j)."""
    return np.array([(i, j) for i in range(n) for j in range(n) if i != j], dtype=np.int64).reshape(-1, 2)


def upper_tri_pairs(n):
    """hisfrag.py:166-167: torch.combinations(arange(n), r=2, with_replacement=True)."""
    return torch.combinations(torch.arange(n), r=2, with_replacement=True).numpy().astype(np.int64)


def sampler_sizes(indexes, num_replicas):
    """data/samplers.py:108-123 (DistributedIndicatesSampler.__init__ boundary computation)."""
    indexes = torch.as_tensor(indexes)
    n_samples_per_rep = math.ceil(len(indexes) / num_replicas)
    indices = torch.split(indexes, n_samples_per_rep)
    sizes = [0]
    for i in range(1, len(indices)):
        if indices[i][0] == indices[i - 1][-1]:
            sizes.append(indices[i][0].item() - 1)
        else:
            sizes.append(indices[i][0].item())
    sizes.append(indexes[-1].item() + 1)
    return sizes


def crop_geometry(img_h, img_w, piece_width, erosion):
    """paikin_tal_solver/puzzle_importer.py:196-213 (grid), :224 (eroded side), :430-446 (centre_crop offset)."""
    numb_cols = int(math.floor(img_w / piece_width))
    numb_rows = int(math.floor(img_h / piece_width))
    top = (img_h - numb_rows * piece_width) // 2
    left = (img_w - numb_cols * piece_width) // 2
    side = math.ceil(piece_width * (1 - erosion))
    crop = side if side < piece_width else piece_width
    off = int(round((piece_width - crop) / 2.0))
    return numb_rows, numb_cols, top, left, crop, off


def puzzle_distance_lookup(logits, i, j, side_i, side_j):
    """evaluation.py:109-131: pred = 1 - sigmoid(logits[i, j]); side codes: 0 top, 1 right, 2 bottom, 3 left
    (paikin_tal_solver/puzzle_piece.py PuzzlePieceSide values)."""
    pred = 1.0 - torch.sigmoid(torch.as_tensor(logits[i][j], dtype=torch.float32)).numpy()
    top, right, bottom, left = 0, 1, 2, 3
    if side_j == left and side_i == right:
        return pred[0] * 1000.
    if side_j == right and side_i == left:
        return pred[2] * 1000.
    if side_j == top and side_i == bottom:
        return pred[1] * 1000.
    if side_j == bottom and side_i == top:
        return pred[3] * 1000.
    return float('inf')


# ---------------------------------------------------------------------------------------------------------------
# retrieval metrics of the Hisfrag consumer (misc/wi19_evaluate.py:12-56) -- used to check the north-star criterion
# "identical retrieval top-1 / mAP to 3 decimals" on score matrices produced by the CUDA path and by this oracle
# ---------------------------------------------------------------------------------------------------------------
def wi19_metrics(distance_matrix, labels, kind=None):
    """get_metrics(distance_matrix, labels, remove_self_column=True) of misc/wi19_evaluate.py:12-22 (``kind`` is passed
    to argsort: None = numpy's default as the reference calls it -- its order among TIED distances is implementation
    defined; 'stable' = ties by ascending index, the convention of the device evaluator):
    rows sorted by ascending distance (numpy argsort, :28), the first column (self) dropped (:29-30), relevance =
    same label (:26-27); mAP over non-singleton queries of mean(precision@rank over relevant ranks) (:49-56),
    top-1 (:18), Pr@10 / Pr@100 (:7-9). Returns (mAP, top_1, pr_a_k10, pr_a_k100)."""
    import numpy as np
    D = np.asarray(distance_matrix)
    classes = np.asarray(labels)
    correct = classes[None, :] == classes[:, None]
    order = np.argsort(D, axis=1, kind=kind)[:, 1:]
    rel = correct[np.arange(order.shape[0], dtype='int64')[:, None], order]
    ranks = np.cumsum(np.ones_like(rel), axis=1)
    precision_at = np.cumsum(rel, axis=1).astype('float') / ranks
    keep = rel.sum(axis=1) > 0
    ap = (precision_at[keep] * rel[keep]).sum(axis=1) / rel[keep].sum(axis=1)
    m_ap = ap.mean()
    top_1 = rel[:, 0].sum() / len(rel)

    def pr_at(k):
        with np.errstate(invalid='ignore', divide='ignore'):   # a singleton query gives 0 / 0 = nan, as in the reference
            v = rel[:, :k].sum(axis=1) / np.minimum(rel.sum(axis=1), k)
        return v.sum() / len(v)

    return float(m_ap), float(top_1), float(pr_at(10)), float(pr_at(100))


def sim_to_distance(sim):
    """hisfrag.py:283-296: the similarity logits stored as fp16, then ``1 - similarity`` in fp16."""
    return (1 - torch.as_tensor(sim, dtype=torch.float32).type(torch.float16)).numpy()


def wi19_rows(distance_matrix, labels):
    """Per query row, with ties by ascending index: (relevant items left after the first sorted column is dropped,
    sum over them of precision at their rank, top-1 hit, hits within 10, hits within 100) -- the integers / sums the
    device evaluator returns, from which get_metrics' four numbers follow."""
    import numpy as np
    D = np.asarray(distance_matrix)
    classes = np.asarray(labels)
    order = np.argsort(D, axis=1, kind='stable')[:, 1:]
    rel = (classes[None, :] == classes[:, None])[np.arange(len(D))[:, None], order]
    prec = np.cumsum(rel, axis=1) / np.arange(1, rel.shape[1] + 1)
    return (rel.sum(1).astype(np.int32), (prec * rel).sum(1), rel[:, 0].astype(np.int32),
            rel[:, :10].sum(1).astype(np.int32), rel[:, :100].sum(1).astype(np.int32))


# ---------------------------------------------------------------------------------------------------------------
# distance tables of the puzzle solver: what InterPieceDistance.__init__ computes from the
# distance_function callback of evaluation.py:116-131 (paikin_tal_solver/inter_piece_distance.py:437-475). Pinned to
# the reference's own class by tests/golden/solver_tables.npz (tests/golden/make_golden_tables.py).
# ---------------------------------------------------------------------------------------------------------------
SIDE_BIN = (3, 0, 1, 2)   # PuzzlePieceSide value (top 0, right 1, bottom 2, left 3) -> score bin (evaluation.py:118-129)
_MAXSIZE = 2 ** 63 - 1    # sys.maxsize on the 64-bit CPython the reference runs on


def solver_tables(distance, order, scalar_rules='numpy1'):
    """scalar_rules: 'numpy1' = ``pred[k] * 1000.`` promotes to float64 (NumPy 1.x, the reference's pinned stack);
    'numpy2' = stays float32 (NEP 50; what the reference code computes in this container's NumPy 2.3).
    distance [N, N, 4] fp32 = 1 - sigmoid(logits), indexed by origin piece id; order[k] = origin id of the piece at
    list position k (evaluation.py:87 shuffles the list; InterPieceDistance numbers pieces by position, :437-441).
    Type-1 puzzles: the only valid neighbour side is the complementary one (:816-819), so tables are [N, 4, N].
    Returns the dict of arrays tests/golden/make_golden_tables.py stores."""
    import numpy as np
    d = np.asarray(distance, dtype=np.float32)
    order = np.asarray(order)
    n = len(order)
    asym = np.full((n, 4, n), 2 ** 31 - 1, dtype=np.uint32)                   # :204-207
    compat = np.full((n, 4, n), np.inf, dtype=np.float32)                     # :340-343
    mutual = np.full((n, 4, n), np.inf, dtype=np.float32)                     # :369-372
    min_d = np.zeros((n, 4), dtype=np.int64)
    second_d = np.zeros((n, 4), dtype=np.int64)
    cand = np.zeros((n, 4, n), dtype=bool)
    for i in range(n):
        for s in range(4):
            mn, sec = _MAXSIZE - 1, _MAXSIZE                                   # :283-288
            cands = []
            for j in range(n):
                if j == i:
                    continue
                # evaluation.py:118-129: float32 element * python float (float64 under NumPy 1.x, float32 under
                # NumPy >= 2), then the uint32 store of :229 truncates
                v = d[order[i], order[j], SIDE_BIN[s]]
                if scalar_rules == 'numpy2':
                    dist = int(np.uint32(np.float32(v) * np.float32(1000.)))
                else:
                    dist = int(np.uint32(np.float64(v) * np.float64(1000.)))
                asym[i, s, j] = dist
                if dist < mn:                                                  # :256-272
                    sec, mn, cands = mn, dist, [j]
                elif dist == mn:
                    sec = dist
                    cands.append(j)
                elif dist < sec:
                    sec = dist
            min_d[i, s], second_d[i, s] = mn, sec
            cand[i, s, cands] = True
            for j in range(n):                                                 # :345-367
                if j == i:
                    continue
                if asym[i, s, j] == 0:
                    c = 1
                elif sec == 0:
                    c = -_MAXSIZE
                else:
                    c = 1 - 1.0 * float(asym[i, s, j]) / sec
                compat[i, s, j] = c
    for i in range(n):                                                         # :489-524
        for s in range(4):
            cs = (s + 2) % 4
            for j in range(i + 1, n):
                m = np.float32(compat[i, s, j] + compat[j, cs, i]) / np.float32(2)
                mutual[i, s, j] = m
                mutual[j, cs, i] = m
    # best buddies (:626-648) with _ALLOW_MULTIPLE_BEST_BUDDIES = False (:21, :76-84): a tied minimum has no candidate
    single = cand.sum(-1) == 1
    first = cand.argmax(-1)
    bb = np.full((n, 4), -1, dtype=np.int32)
    for i in range(n):
        for s in range(4):
            cs = (s + 2) % 4
            if single[i, s]:
                j = int(first[i, s])
                if single[j, cs] and first[j, cs] == i:
                    bb[i, s] = j
    # starter piece ordering (:650-719): 4 x own best buddies + best buddies of those, then summed mutual compatibility
    info = []
    for i in range(n):
        ids, total = [], 0
        for s in range(4):
            if bb[i, s] >= 0:
                ids.append(int(bb[i, s]))
                total = total + mutual[i, s, bb[i, s]]
        info.append((ids, total))
    ordering = []
    for i in range(n):
        numb = 4 * len(info[i][0]) + sum(len(info[b][0]) for b in info[i][0])
        ordering.append((i, numb, info[i][1]))
    ordering.sort(key=lambda t: (t[1], t[2]), reverse=True)
    return {
        'asym_dist': asym, 'asym_compat': compat, 'mutual_compat': mutual, 'min_dist': min_d, 'second_dist': second_d,
        'candidates': cand, 'best_buddy': bb,
        'start_order': np.array([(a, b) for (a, b, _) in ordering], dtype=np.int64).reshape(-1, 2),
        'start_compat': np.array([float(c) for (_, _, c) in ordering], dtype=np.float64),
    }


# ---------------------------------------------------------------------------------------------------------------
# piece preparation: what PiecesDataset.__getitem__ + TwoImgSyncEval do to each piece
# (data/datasets/pieces_dataset.py:34-56, data/transforms.py:12-26). `prepare_pieces` makes the reference's own
# library calls (cv2, PIL via torchvision -- all present on the GPU box too); `lab2rgb_u8` / `pil_resize_bilinear_u8`
# restate the integer arithmetic of those libraries in numpy and are pinned to them by tests/test_piece_prep.py.
# ---------------------------------------------------------------------------------------------------------------
def prepare_pieces(img_bgr, piece_width, erosion, img_size):
    """BGR uint8 [H, W, 3] -> fp32 [N, 3, S, S] in piece-id order, through the calls the reference makes:
    cv2 BGR2LAB on the image (puzzle_importer.py:136-156), centred grid + eroded centre crop (:196-232, :430-446),
    then per piece cv2 LAB2RGB -> ToPILImage -> Resize(S) (PIL bilinear) -> ToTensor -> Normalize(.5, .5)."""
    import cv2
    from torchvision import transforms
    lab = cv2.cvtColor(img_bgr, cv2.COLOR_BGR2LAB)
    rows, cols, top, left, side, off = crop_geometry(img_bgr.shape[0], img_bgr.shape[1], piece_width, erosion)
    tf = transforms.Compose([transforms.ToPILImage(), transforms.Resize(img_size), transforms.ToTensor(),
                             transforms.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))])
    out = []
    for r in range(rows):
        for c in range(cols):
            y0, x0 = top + r * piece_width + off, left + c * piece_width + off
            piece = np.ascontiguousarray(lab[y0:y0 + side, x0:x0 + side])
            out.append(tf(cv2.cvtColor(piece, cv2.COLOR_LAB2RGB)))
    return torch.stack(out)


def lab2rgb_u8(lab):
    """OpenCV's bit-exact 8-bit Lab -> sRGB (imgproc color_lab.cpp, Lab2RGBinteger): L -> (Y, f(Y)) tables in 2^14
    fixed point, a / b offsets by multiply-shift, f^-1 by integer division (toward zero) or integer cube, 3x3 integer
    matrix descaled to a 12-bit index into the inverse-gamma table. Tables: tools/gen_lab_tables.py."""
    import importlib.util
    import os
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('gen_lab_tables', os.path.join(here, 'tools', 'gen_lab_tables.py'))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    y_tab, fy_tab = gen.lab_to_y_fy()
    gamma = _inv_gamma_table(os.path.join(here, 'vit-ed_b200', 'csrc', 'lab_tables.inc'))
    lab = np.asarray(lab).astype(np.int64)
    L, a, b = lab[..., 0], lab[..., 1], lab[..., 2]
    base = 1 << 14
    adiv = ((5 * a * 53687 + (1 << 7)) >> 13) - 128 * base // 500
    bdiv = ((b * 41943 + (1 << 4)) >> 9) - 128 * base // 200 + 1
    x, z, y = gen.ab_to_xz(fy_tab[L] + adiv), gen.ab_to_xz(fy_tab[L] - bdiv), y_tab[L]
    out = [gamma[np.clip((gen.COEFFS[c, 0] * x + gen.COEFFS[c, 1] * y + gen.COEFFS[c, 2] * z + (1 << 13)) >> 14, 0, 4095)]
           for c in range(3)]
    return np.stack(out, -1).astype(np.uint8)


def _inv_gamma_table(path):
    text = open(path).read()
    body = text[text.index('kInvGamma[4096]'):]
    body = body[body.index('{') + 1:body.index('}')]
    return np.array([int(v) for v in body.replace('\n', ' ').split(',') if v.strip()], dtype=np.int64)


def pil_bilinear_coeffs(in_size, out_size):
    """Pillow Resample.c precompute_coeffs + normalize_coeffs_8bpc, bilinear filter (support 1), whole-axis box."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    bounds, coefs = [], []
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = [max(0.0, 1.0 - abs((x + xmin - center + 0.5) * ss)) for x in range(xmax)]
        ww = sum(k)
        k = [v / ww if ww != 0.0 else v for v in k]
        coefs.append([int(0.5 + v * (1 << 22)) for v in k])
        bounds.append((xmin, xmax))
    return bounds, coefs


def pil_resize_bilinear_u8(img, out_size):
    """Pillow's 8-bit two-pass resize of a square [s, s, C] uint8 image to [S, S, C]: horizontal pass, uint8 rounding
    (clip8 of the 22-bit fixed-point sum), vertical pass (ImagingResampleHorizontal_8bpc / Vertical_8bpc)."""
    img = np.asarray(img).astype(np.int64)
    s = img.shape[0]
    bounds, coefs = pil_bilinear_coeffs(s, out_size)

    def one_pass(a):   # along axis 1
        out = np.zeros((a.shape[0], out_size, a.shape[2]), dtype=np.int64)
        for xx, ((xmin, n), k) in enumerate(zip(bounds, coefs)):
            acc = np.full((a.shape[0], a.shape[2]), 1 << 21, dtype=np.int64)
            for t in range(n):
                acc += a[:, xmin + t] * k[t]
            out[:, xx] = np.clip(acc >> 22, 0, 255)
        return out

    h = one_pass(img)
    return one_pass(h.transpose(1, 0, 2)).transpose(1, 0, 2).astype(np.uint8)


def prepare_pieces_restated(img_lab, piece_width, erosion, img_size):
    """The same result as `prepare_pieces` from the LAB image, through the numpy restatements only (fp32 division by
    255, then (x - 0.5) / 0.5 as torch does)."""
    rows, cols, top, left, side, off = crop_geometry(img_lab.shape[0], img_lab.shape[1], piece_width, erosion)
    out = []
    for r in range(rows):
        for c in range(cols):
            y0, x0 = top + r * piece_width + off, left + c * piece_width + off
            rgb = pil_resize_bilinear_u8(lab2rgb_u8(img_lab[y0:y0 + side, x0:x0 + side]), img_size)
            v = rgb.astype(np.float32) / np.float32(255)
            out.append(((v - np.float32(0.5)) / np.float32(0.5)).transpose(2, 0, 1))
    return torch.from_numpy(np.stack(out))


# ---------------------------------------------------------------------------------------------------------------
# Hisfrag training step (product side not built): pair construction of prepare_data
# (hisfrag.py:117-155) and loss / parameter gradients of one step (train_step :157-159, BCEWithLogitsLoss :60-61) by
# autograd over the functional model above. Pinned by tests/golden/train_step.npz, which the reference's own
# prepare_data and model file wrote (tests/golden/make_golden_train.py).
# ---------------------------------------------------------------------------------------------------------------
def train_pairs(targets, perm_seed):
    """(groups [P, 2], labels [P, 1]): for every i the later items j > i with the same target are positive pairs
    (i, j), those with another target negative pairs (hisfrag.py:121-137); at most twice as many negatives as
    positives are kept, drawn by ``torch.randperm`` (:142-143; seeded here), positives first (:145-147)."""
    t = torch.as_tensor(targets)
    n = len(t)
    pos = [(i, j) for i in range(n) for j in range(i, n) if j != i and t[j] == t[i]]
    neg = [(i, j) for i in range(n) for j in range(i, n) if t[j] != t[i]]
    pos_g = torch.tensor(pos, dtype=torch.long).view(-1, 2)
    neg_g = torch.tensor(neg, dtype=torch.long).view(-1, 2)
    neg_length = min(len(neg_g), int(2 * len(pos_g)))
    torch.manual_seed(perm_seed)
    neg_g = neg_g[torch.randperm(len(neg_g))[:neg_length]]
    labels = torch.tensor([1.] * len(pos_g) + [0.] * len(neg_g), dtype=torch.float32).view(-1, 1)
    return torch.cat([pos_g, neg_g], dim=0), labels


def train_step(sd, num_heads, samples, targets, perm_seed):
    """One training step without dropout / stochastic depth: encode the batch (with gradient, :149-150), decode
    ``model(tokens[groups[:, 1]], samples[groups[:, 0]])`` (:152-154, :157-159), BCE-with-logits against the pair
    labels. Returns (loss, logits, {parameter name: gradient})."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    groups, labels = train_pairs(targets, perm_seed)
    with torch.enable_grad():
        tokens = forward_first_part(samples, params, num_heads)
        x2 = forward_second_part(tokens[groups[:, 1]], samples[groups[:, 0]], params, num_heads)
        logits = forward_head(x2, params)
        loss = F.binary_cross_entropy_with_logits(logits, labels)
        loss.backward()
    return loss.detach(), logits.detach(), groups, labels, {k: p.grad for k, p in params.items()}
