"""The reference experiment configs this path is quoted on, as attribute trees shaped like the yacs nodes that
``build_model(config)`` reads (reference config.py:52-79 defaults + configs/*/*.yaml). yacs itself is not needed."""
from types import SimpleNamespace as _NS


def _cfg(name, img_size, num_classes, patch, embed=768, depth=8, c_depth=8, heads=12, dataset='', erosion=0.07):
    # defaults: config.py:68-79 (PATCH 16, EMBED 768, DEPTH 8, C_DEPTH 8, HEADS 12, MLP 4, QKV_BIAS True)
    return _NS(
        MODEL=_NS(TYPE='pjs', NAME=name, NUM_CLASSES=num_classes, DROP_PATH_RATE=0.1,
                  PJS=_NS(PATCH_SIZE=patch, IN_CHANS=3, EMBED_DIM=embed, DEPTH=depth, C_DEPTH=c_depth, NUM_HEADS=heads,
                          MLP_RATIO=4., QKV_BIAS=True, QK_SCALE=None, KEEP_ATTN=False, ARCH_VERSION='v1')),
        DATA=_NS(IMG_SIZE=img_size, DATASET=dataset, EROSION_RATIO=erosion, BATCH_SIZE=128, TEST_BATCH_SIZE=512),
        AMP_ENABLE=True, SEED=0,
    )


def puzzle_patch8_64():
    """configs/puzzle/div2k_erosion7_4bin_patch8_64.yaml:1-11"""
    return _cfg('div2k_erosion7_4bin_patch8_64', 64, 4, 8, embed=384, dataset='div2k')


def hisfrag20_patch16_512():
    """configs/hisfrag/hisfrag20_patch16_512.yaml:1-14"""
    return _cfg('hisfrag20_patch16_512', 512, 1, 16, embed=384, depth=12, c_depth=12, heads=6, dataset='hisfrag20')


def test_patch32_64():
    """configs/test/test_pjs_hisfrag20_patch32_64.yaml:1-14 (the reference's tiny smoke shape)"""
    return _cfg('test_hisfrag20', 64, 1, 32, embed=32, depth=1, c_depth=1, heads=1, dataset='hisfrag20')


CONFIGS = {
    'puzzle': puzzle_patch8_64,
    'hisfrag': hisfrag20_patch16_512,
    'test': test_patch32_64,
}


def get_config(name):
    return CONFIGS[name]()
