"""B200-native (sm_100a) implementation of the OutGridViT `OutGridBlock` hot path
(Outlooker -> MBConv -> GridAttn -> MLP, forward + backward) behind the reference's nn.Module
surface.  All compute runs in libogvit.so (hand-written CUDA, C-ABI in include/ogv.h); there is no
CPU path and no Triton / torch.compile path.
"""
from .dropin import find_reference_root, install, uninstall
from .config import BASELINE_CONFIGS, StageCfg, build_stages, load_yaml
from .functional import cross_entropy
from .model import (ConvStem, Downsample, DownsampleConfig, MaxOutNet, OutlookerFrontGridNet, build_model,
                    make_dpr)
from .modules import (MLP, AttentionConfig, DropPath, GridAttention2D, GridAttention2DConfig, GridOnlyBlock,
                      LayerNorm2d, MBConv, MBConvConfig, MLP2d, MultiHeadSelfAttention, OutGridBlock,
                      OutlookAttention2d, OutlookerBlock2d, SqueezeExcite, grid_partition, grid_unpartition,
                      make_activation)

__version__ = "0.1.0"
