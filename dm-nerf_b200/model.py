"""DM_NeRF network with the reference's constructor, parameter names and shapes
(networks/dm_nerf.py:58-106) so checkpoints, `Adam(model.parameters())`, `.to(device)`, `.train()/.eval()`
and `print(model)` behave identically; forward() runs the fused CUDA kernel."""
import torch
import torch.nn as nn

from . import _lib
from .engine import get_context


class DM_NeRF(nn.Module):
    def __init__(self, D=8, W=256, input_ch_pts=3, input_ch_views=3, skips=[4], ins_num=None):
        super().__init__()
        self.skips = skips
        self.input_ch_pts = input_ch_pts
        self.input_ch_views = input_ch_views
        self.mlps = nn.ModuleList(
            [nn.Linear(input_ch_pts, W)]
            + [nn.Linear(W + input_ch_pts, W) if i in skips else nn.Linear(W, W) for i in range(D - 1)])
        self.rgb_feature_linear = nn.Linear(W, W)
        self.ins_feature_linear = nn.Linear(W, W)
        self.rgb_feature_linears = nn.ModuleList([nn.Linear(W + input_ch_views, W // 2)])
        self.ins_feature_linears = nn.ModuleList([nn.Linear(W, W // 2)])
        self.density_linear = nn.Linear(W, 1)
        self.ins_linear = nn.Linear(W // 2, ins_num + 1)
        self.rgb_linear = nn.Linear(W // 2, 3)
        self._check_shape(D, W)

    def _check_shape(self, D, W):
        if not (D == 8 and W == 256 and self.input_ch_pts == 63 and self.input_ch_views == 27
                and list(self.skips) == [4]):
            raise NotImplementedError(
                "the B200 kernels are specialised for DM_NeRF(D=8, W=256, input_ch_pts=63, input_ch_views=27, "
                "skips=[4]) -- the only configuration config.create_nerf builds (config.py:126-138)")

    @property
    def ins_num(self):
        return self.ins_linear.out_features - 1

    def forward(self, x, impl=_lib.IMPL_AUTO):
        from .autograd import mlp_forward
        return mlp_forward(self, x, impl)
