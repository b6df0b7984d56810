"""GPU: max |logit - reference fixture| per golden model case (tests/golden/model_*.npz were produced by the
reference's own model file in fp32), for the library in use (fp16 default, or VITED_LIB=... for the bf16 build)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import vited_b200  # noqa: E402
from tests import helpers  # noqa: E402

for name in helpers.MODEL_CASES:
    z, kw = helpers.load_model_case(name)
    model, _ = helpers.make_gpu_model(kw, int(z['weight_seed']))
    x1, x2 = helpers.case_inputs(z, kw)
    x1, x2 = x1.cuda(), x2.cuda()
    tokens = model(x1, forward_first_part=True)
    two_phase = model(tokens, x2).cpu().numpy()
    tok_err = np.abs(tokens[:, :4].cpu().numpy() - z['tokens_head']).max()
    err = np.abs(two_phase - z['two_phase'])
    print(f'[{vited_b200.ACT_NAME}] {name:24s} logits max err {err.max():.5f} mean {err.mean():.5f} '
          f'(|logit| max {np.abs(z["two_phase"]).max():.3f}); encoder tokens max err {tok_err:.5f}')
