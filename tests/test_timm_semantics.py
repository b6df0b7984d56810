"""CPU: the timm-owned arithmetic of the model (PatchEmbed conv, class token + position embedding, Mlp with exact GELU,
LayerNorm eps 1e-6, the pre-norm block structure, final norm -> token 0 -> head) cross-checked against a SECOND,
independent implementation of the same published architecture: torchvision's VisionTransformer.

Why: timm==0.9.2 (reference requirements.txt:2) is not in the image, so the fixtures that pin the oracle were produced
by running the reference's own model file against tests/golden/_timm_shim.py, a restatement of the five timm symbols
it imports ("parity unpinned" for the timm-owned part). torchvision's ViT is not timm, but it is a
third-party implementation of the identical ViT forward (weights convert 1:1 between the two), it is in the image, and
neither the shim nor the oracle was written from it. Agreement of all three on the same weights rules out a private
misreading of the architecture in the shim / oracle (patch order, where the class token and position embedding go,
GELU flavour, LayerNorm eps, which token reaches the head).
"""
import torch
import torch.nn as nn

from oracle import vited_oracle as orc
from tests.golden import _timm_shim as shim

IMG, P, D, H, HID, C = 32, 8, 96, 3, 384, 4


def _tv_vit(num_layers, seed):
    from torchvision.models.vision_transformer import VisionTransformer
    torch.manual_seed(seed)
    tv = VisionTransformer(image_size=IMG, patch_size=P, num_layers=num_layers, num_heads=H, hidden_dim=D, mlp_dim=HID,
                           num_classes=C).eval()
    with torch.no_grad():   # torchvision zero-initialises several tensors; make every parameter matter
        for name, p in tv.named_parameters():
            if 'ln' in name and name.endswith('weight'):
                p.copy_(1 + 0.2 * torch.randn_like(p))
            else:
                p.copy_(0.1 * torch.randn_like(p))
    return tv


def _state_dict_from_tv(tv):
    """torchvision parameter names -> the reference's state_dict keys (encoder part + head)."""
    sd = {'patch_embed.proj.weight': tv.conv_proj.weight, 'patch_embed.proj.bias': tv.conv_proj.bias,
          'cls_token': tv.class_token, 'pos_embed': tv.encoder.pos_embedding,
          'norm.weight': tv.encoder.ln.weight, 'norm.bias': tv.encoder.ln.bias,
          'head.weight': tv.heads.head.weight, 'head.bias': tv.heads.head.bias}
    for l, blk in enumerate(tv.encoder.layers):
        b = f'blocks.{l}'
        sd[b + '.norm1.weight'], sd[b + '.norm1.bias'] = blk.ln_1.weight, blk.ln_1.bias
        sd[b + '.attn.qkv.weight'], sd[b + '.attn.qkv.bias'] = blk.self_attention.in_proj_weight, blk.self_attention.in_proj_bias
        sd[b + '.attn.proj.weight'], sd[b + '.attn.proj.bias'] = blk.self_attention.out_proj.weight, blk.self_attention.out_proj.bias
        sd[b + '.norm2.weight'], sd[b + '.norm2.bias'] = blk.ln_2.weight, blk.ln_2.bias
        sd[b + '.mlp.fc1.weight'], sd[b + '.mlp.fc1.bias'] = blk.mlp[0].weight, blk.mlp[0].bias
        sd[b + '.mlp.fc2.weight'], sd[b + '.mlp.fc2.bias'] = blk.mlp[3].weight, blk.mlp[3].bias
    return {k: v.detach().clone() for k, v in sd.items()}


@torch.no_grad()
def test_oracle_vit_pipeline_equals_torchvision_vit():
    """prepare_x2 (PatchEmbed, class token, position embedding) -> pre-norm blocks (attention, Mlp / GELU) -> final
    LayerNorm -> token 0 -> head, composed from the oracle's functions, equals torchvision's ViT on the same weights."""
    tv = _tv_vit(num_layers=3, seed=1)
    sd = _state_dict_from_tv(tv)
    images = torch.randn(5, 3, IMG, IMG)
    x = orc.prepare_x2(images, sd)
    for l in range(3):
        x = orc.block(x, sd, f'blocks.{l}', H)
    got = orc.forward_head(orc._ln(x, sd, 'norm'), sd)
    want = tv(images)
    assert got.shape == want.shape == (5, C)
    assert (got - want).abs().max().item() < 2e-5 * max(1.0, want.abs().max().item())
    # patch tokens alone (what forward_first_part starts from): same patch order and in-patch index order
    tok = tv._process_input(images)
    assert (orc.patch_embed(images, sd) - tok).abs().max().item() < 1e-5
    # the comparison has teeth: approximate GELU or the default LayerNorm eps would not pass at this tolerance
    x = orc.prepare_x2(images, sd)
    y = x + orc.attention(orc._ln(x, sd, 'blocks.0.norm1'), sd, 'blocks.0.attn', H)
    h = orc._ln(y, sd, 'blocks.0.norm2')
    exact = orc.mlp(h, sd, 'blocks.0.mlp')
    tanh = torch.nn.functional.linear(torch.nn.functional.gelu(
        torch.nn.functional.linear(h, sd['blocks.0.mlp.fc1.weight'], sd['blocks.0.mlp.fc1.bias']), approximate='tanh'),
        sd['blocks.0.mlp.fc2.weight'], sd['blocks.0.mlp.fc2.bias'])
    assert (exact - tanh).abs().max().item() > 1e-4


@torch.no_grad()
def test_timm_shim_symbols_equal_torchvision_counterparts():
    """The stand-ins the golden generator ran the reference model file against: PatchEmbed, _pos_embed, LayerNorm eps,
    forward_head (as a depth-0 VisionTransformer) and Mlp, each against torchvision's implementation."""
    tv = _tv_vit(num_layers=1, seed=2)
    images = torch.randn(4, 3, IMG, IMG)
    vit = shim.VisionTransformer(img_size=IMG, patch_size=P, in_chans=3, num_classes=C, embed_dim=D, depth=0,
                                 num_heads=H).eval()
    vit.patch_embed.proj.load_state_dict(tv.conv_proj.state_dict())
    vit.cls_token.copy_(tv.class_token)
    vit.pos_embed.copy_(tv.encoder.pos_embedding)
    vit.norm.load_state_dict(tv.encoder.ln.state_dict())
    vit.head.load_state_dict(tv.heads.head.state_dict())
    assert vit.norm.eps == tv.encoder.ln.eps == 1e-6
    tok = vit.patch_embed(images)
    assert torch.allclose(tok, tv._process_input(images), atol=1e-6)
    x = vit._pos_embed(tok)
    want_x = torch.cat([tv.class_token.expand(4, -1, -1), tv._process_input(images)], dim=1) + tv.encoder.pos_embedding
    assert torch.allclose(x, want_x, atol=1e-6)
    # no blocks in between: final norm -> token 0 -> head, against the same steps of torchvision's forward
    assert torch.allclose(vit.forward_head(vit.norm(x)), tv.heads(tv.encoder.ln(want_x)[:, 0]), atol=1e-5)
    # Mlp = fc1 -> exact GELU -> fc2 against torchvision's MLPBlock
    blk = tv.encoder.layers[0]
    m = shim.Mlp(in_features=D, hidden_features=HID).eval()
    m.fc1.load_state_dict(blk.mlp[0].state_dict())
    m.fc2.load_state_dict(blk.mlp[3].state_dict())
    h = torch.randn(4, 17, D)
    assert torch.allclose(m(h), blk.mlp(h), atol=1e-6)
    assert isinstance(m.act, nn.GELU) and m.act.approximate == 'none'
