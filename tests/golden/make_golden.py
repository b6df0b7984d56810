"""Generates tests/golden/*.npz by RUNNING THE REFERENCE in the build container (where /root/reference exists).

  python tests/golden/make_golden.py

* model vectors: the reference's own models/vision_transformer.py (imported from /root/reference, timm provided by
  _timm_shim.py) evaluated in fp32 on CPU on seeded synthetic weights/inputs (vited_b200.synthetic);
* integer vectors: the reference's own PiecesDataset, Puzzle (erosion crop), DistributedIndicatesSampler,
  torch.combinations pair set and the evaluation.py distance closure semantics.
The fixtures are small and committed; nothing at test time reads /root/reference.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import _timm_shim  # noqa: E402
import vited_b200  # noqa: E402
from vited_b200 import synthetic  # noqa: E402

CASES = {
    # name: (ctor kwargs, n_pairs, weight seed, image seed)
    'test_patch32_64': (dict(img_size=64, patch_size=32, num_classes=1, embed_dim=32, depth=1, c_depth=1, num_heads=1), 6, 1, 11),
    'small_hd64': (dict(img_size=64, patch_size=16, num_classes=1, embed_dim=128, depth=2, c_depth=2, num_heads=2), 5, 2, 12),
    'small_hd32': (dict(img_size=32, patch_size=8, num_classes=4, embed_dim=96, depth=2, c_depth=3, num_heads=3), 7, 3, 13),
    'puzzle_patch8_64': (dict(img_size=64, patch_size=8, num_classes=4, embed_dim=384, depth=8, c_depth=8, num_heads=12), 6, 0, 14),
    'hisfrag20_patch16_512': (dict(img_size=512, patch_size=16, num_classes=1, embed_dim=384, depth=12, c_depth=12, num_heads=6), 2, 0, 15),
}


def load_reference_model_module():
    _timm_shim.install()
    spec = importlib.util.spec_from_file_location('ref_vision_transformer', os.path.join(REF, 'models', 'vision_transformer.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def model_vectors(ref_mod):
    for name, (kw, n_pairs, wseed, iseed) in CASES.items():
        torch.manual_seed(0)
        model = ref_mod.VisionTransformerCustom(mlp_ratio=4., qkv_bias=True, **kw).eval()
        sd = synthetic.synthetic_state_dict(model, seed=wseed)
        missing, unexpected = model.load_state_dict(sd, strict=True), None
        imgs = synthetic.synthetic_images(2 * n_pairs, kw['img_size'], seed=iseed)
        x1, x2 = imgs[:n_pairs], imgs[n_pairs:]
        with torch.no_grad():
            one_shot = model(torch.stack([x1, x2], dim=1))
            tokens = model(x1, forward_first_part=True)
            two_phase = model(tokens, x2)
        out = dict(
            kwargs=np.array(repr(kw)), n_pairs=n_pairs, weight_seed=wseed, image_seed=iseed,
            one_shot=one_shot.numpy(), two_phase=two_phase.numpy(),
            tokens_head=tokens[:, :4, :].numpy(),            # first 4 tokens of every item
            tokens_sum=tokens.double().sum(dim=(1, 2)).numpy(),
            tokens_abs_sum=tokens.double().abs().sum(dim=(1, 2)).numpy(),
            n_state_dict=len(sd), n_params=sum(p.numel() for p in model.parameters()),
            keys=np.array(sorted(sd.keys())),
        )
        np.savez_compressed(os.path.join(HERE, f'model_{name}.npz'), **out)
        print(name, 'one-shot vs two-phase max abs diff', float((one_shot - two_phase).abs().max()),
              'logits', one_shot.flatten()[:4].tolist())


def integer_vectors():
    sys.path.insert(0, REF)
    out = {}
    # ---- pair enumeration: the reference dataset class itself
    from data.datasets.pieces_dataset import PiecesDataset
    for n in (1, 2, 5, 9):
        ds = PiecesDataset([object()] * n)
        out[f'entries_{n}'] = np.array(ds.entries, dtype=np.int64).reshape(-1, 2)
    # ---- hisfrag pair set
    for n in (1, 4, 7):
        out[f'combos_{n}'] = torch.combinations(torch.arange(n).type(torch.int), r=2, with_replacement=True).numpy()
    # ---- sampler boundaries: the reference sampler class (pytorch_metric_learning stubbed: only imported, not used)
    pml = types.ModuleType('pytorch_metric_learning')
    pml_utils = types.ModuleType('pytorch_metric_learning.utils')
    pml_cf = types.ModuleType('pytorch_metric_learning.utils.common_functions')
    pml_cf.safe_random_choice = lambda *a, **k: None
    pml_cf.NUMPY_RANDOM = np.random
    pml.utils, pml_utils.common_functions = pml_utils, pml_cf
    sys.modules.update({'pytorch_metric_learning': pml, 'pytorch_metric_learning.utils': pml_utils,
                        'pytorch_metric_learning.utils.common_functions': pml_cf})
    # torch>=2.2 removed Sampler.__init__(data_source); the reference pins torch~=2.1 and calls super().__init__(None)
    import torch.utils.data.sampler as _tsampler
    _tsampler.Sampler.__init__ = lambda self, data_source=None: None
    spec = importlib.util.spec_from_file_location('ref_samplers', os.path.join(REF, 'data', 'samplers.py'))
    samplers = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(samplers)
    cases = []
    for n, world in ((16, 2), (16, 4), (33, 8), (100, 8), (257, 4), (4096, 8), (4096, 2)):
        pairs = torch.combinations(torch.arange(n).type(torch.int), r=2, with_replacement=True)
        rows = []
        n_chunks = None
        for rank in range(world):
            try:
                s = samplers.DistributedIndicatesSampler(pairs[:, 0], num_replicas=world, rank=rank)
                rows.append([int(s.samples[0]) if len(s.samples) else -1, int(s.samples[-1]) + 1 if len(s.samples) else -1])
            except IndexError:
                rows.append([-2, -2])
        cases.append((n, world))
        out[f'sampler_{n}_{world}'] = np.array(rows, dtype=np.int64)
    out['sampler_cases'] = np.array(cases, dtype=np.int64)
    # ---- erosion crop + per-piece transform: the reference Puzzle / PiecesDataset / TwoImgSyncEval
    import cv2
    from paikin_tal_solver.puzzle_importer import Puzzle
    Puzzle.print_debug_messages = False
    alb = types.ModuleType('albumentations')
    sys.modules['albumentations'] = alb
    utils_stub = types.ModuleType('misc.utils')
    class UnableToCrop(Exception):
        pass
    utils_stub.UnableToCrop = UnableToCrop
    misc_pkg = types.ModuleType('misc')
    misc_pkg.utils = utils_stub
    sys.modules['misc'] = misc_pkg
    sys.modules['misc.utils'] = utils_stub
    spec = importlib.util.spec_from_file_location('ref_transforms', os.path.join(REF, 'data', 'transforms.py'))
    transforms = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(transforms)
    img = synthetic.synthetic_puzzle_image(3, 4, piece=64, seed=5)
    img = np.pad(img, ((3, 4), (5, 8), (0, 0)), mode='edge')       # not a multiple of 64: exercises the centring
    path = '/tmp/vited_golden_puzzle.png'
    cv2.imwrite(path, img)
    out['puzzle_image'] = img
    for erosion in (0.0, 0.07, 0.14):
        puzzle = Puzzle(0, path, 64, starting_piece_id=0, erosion=erosion)
        pieces = puzzle.pieces
        tag = str(erosion).replace('.', 'p')
        out[f'grid_{tag}'] = np.array(puzzle.grid_size, dtype=np.int64)
        out[f'piece_shape_{tag}'] = np.array(pieces[0].lab_image.shape, dtype=np.int64)
        out[f'pieces_lab_{tag}'] = np.stack([p.lab_image for p in pieces])
        ds = PiecesDataset(pieces, transform=transforms.TwoImgSyncEval(64))
        stacked, label = ds[1]        # entry 1 = pair (0, 2)
        out[f'pair_tensor_{tag}'] = stacked.numpy()
        out[f'pair_entry_{tag}'] = np.array(ds.entries[1], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, 'integer_paths.npz'), **out)
    print('integer vectors:', sorted(out.keys()))


if __name__ == '__main__':
    if not os.path.isdir(REF):
        raise SystemExit('/root/reference is not available here; fixtures can only be regenerated in the build container')
    integer_vectors()
    model_vectors(load_reference_model_module())
