import torch
import torch.nn as nn
import torch.nn.functional as F

def flow_warp(x, flow, interpolation="bilinear", padding_mode="zeros", align_corners=True):
    """mmedit 0.12 flow_warp: x (n,c,h,w), flow (n,h,w,2) in pixels (x, y)."""
    _, _, h, w = x.size()
    grid_y, grid_x = torch.meshgrid(torch.arange(0, h), torch.arange(0, w), indexing="ij")
    grid = torch.stack((grid_x, grid_y), 2).type_as(x)
    grid.requires_grad = False
    grid_flow = grid + flow
    gx = 2.0 * grid_flow[:, :, :, 0] / max(w - 1, 1) - 1.0
    gy = 2.0 * grid_flow[:, :, :, 1] / max(h - 1, 1) - 1.0
    grid_flow = torch.stack((gx, gy), dim=3)
    return F.grid_sample(x, grid_flow, mode=interpolation, padding_mode=padding_mode,
                         align_corners=align_corners)

class PixelShufflePack(nn.Module):
    def __init__(self, in_channels, out_channels, scale_factor, upsample_kernel):
        super().__init__()
        self.scale_factor = scale_factor
        self.upsample_conv = nn.Conv2d(in_channels, out_channels * scale_factor * scale_factor,
                                       upsample_kernel, padding=(upsample_kernel - 1) // 2)
    def forward(self, x):
        return F.pixel_shuffle(self.upsample_conv(x), self.scale_factor)
