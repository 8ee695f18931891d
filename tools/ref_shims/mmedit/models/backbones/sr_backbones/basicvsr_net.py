"""Restatement of the mmedit 0.12.0 pieces FLAIR imports (un-vendored dependency)."""
import torch
import torch.nn as nn
import torch.nn.functional as F
from mmedit.models.common import flow_warp


class _ConvModule(nn.Module):
    def __init__(self, cin, cout, k, s, p, act=True):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, s, p)
        self.activate = nn.ReLU(inplace=True) if act else None
    def forward(self, x):
        x = self.conv(x)
        return self.activate(x) if self.activate is not None else x


class SPyNetBasicModule(nn.Module):
    def __init__(self):
        super().__init__()
        self.basic_module = nn.Sequential(
            _ConvModule(8, 32, 7, 1, 3), _ConvModule(32, 64, 7, 1, 3), _ConvModule(64, 32, 7, 1, 3),
            _ConvModule(32, 16, 7, 1, 3), _ConvModule(16, 2, 7, 1, 3, act=False))
    def forward(self, x):
        return self.basic_module(x)


class SPyNet(nn.Module):
    def __init__(self, pretrained=None):
        super().__init__()
        self.basic_module = nn.ModuleList([SPyNetBasicModule() for _ in range(6)])
        self.register_buffer("mean", torch.Tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1))
        self.register_buffer("std", torch.Tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1))

    def compute_flow(self, ref, supp):
        n, _, h, w = ref.size()
        ref = [(ref - self.mean) / self.std]
        supp = [(supp - self.mean) / self.std]
        for _ in range(5):
            ref.append(F.avg_pool2d(ref[-1], kernel_size=2, stride=2, count_include_pad=False))
            supp.append(F.avg_pool2d(supp[-1], kernel_size=2, stride=2, count_include_pad=False))
        ref, supp = ref[::-1], supp[::-1]
        flow = ref[0].new_zeros(n, 2, h // 32, w // 32)
        for level in range(len(ref)):
            if level == 0:
                flow_up = flow
            else:
                flow_up = F.interpolate(flow, scale_factor=2, mode="bilinear", align_corners=True) * 2.0
            flow = flow_up + self.basic_module[level](torch.cat(
                [ref[level], flow_warp(supp[level], flow_up.permute(0, 2, 3, 1), padding_mode="border"),
                 flow_up], 1))
        return flow

    def forward(self, ref, supp):
        h, w = ref.shape[2:4]
        w_up = w if (w % 32) == 0 else 32 * (w // 32 + 1)
        h_up = h if (h % 32) == 0 else 32 * (h // 32 + 1)
        ref = F.interpolate(ref, size=(h_up, w_up), mode="bilinear", align_corners=False)
        supp = F.interpolate(supp, size=(h_up, w_up), mode="bilinear", align_corners=False)
        flow = F.interpolate(self.compute_flow(ref, supp), size=(h, w), mode="bilinear", align_corners=False)
        flow[:, 0, :, :] *= float(w) / float(w_up)
        flow[:, 1, :, :] *= float(h) / float(h_up)
        return flow


class ResidualBlockNoBN(nn.Module):
    def __init__(self, mid_channels=64, res_scale=1.0):
        super().__init__()
        self.res_scale = res_scale
        self.conv1 = nn.Conv2d(mid_channels, mid_channels, 3, 1, 1, bias=True)
        self.conv2 = nn.Conv2d(mid_channels, mid_channels, 3, 1, 1, bias=True)
        self.relu = nn.ReLU(inplace=True)
        if res_scale == 1.0:
            for m in (self.conv1, self.conv2):
                nn.init.kaiming_normal_(m.weight, a=0, mode="fan_in", nonlinearity="relu")
                m.weight.data *= 0.1
                nn.init.constant_(m.bias, 0)
    def forward(self, x):
        return x + self.conv2(self.relu(self.conv1(x))) * self.res_scale


class ResidualBlocksWithInputConv(nn.Module):
    def __init__(self, in_channels, out_channels=64, num_blocks=30):
        super().__init__()
        main = [nn.Conv2d(in_channels, out_channels, 3, 1, 1, bias=True),
                nn.LeakyReLU(negative_slope=0.1, inplace=True),
                nn.Sequential(*[ResidualBlockNoBN(mid_channels=out_channels) for _ in range(num_blocks)])]
        self.main = nn.Sequential(*main)
    def forward(self, feat):
        return self.main(feat)
