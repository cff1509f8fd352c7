"""PositionNet parameter container (reference GLIGEN/ldm/modules/diffusionmodules/text_grounding_net.py:9-43).
state_dict: linears.{0,2,4}.{weight,bias}, null_positive_feature [in_dim], null_position_feature [4*2*fourier_freqs].
The Fourier embedding + MLP run in the library once per image (ltt_set_conditioning)."""
import torch
import torch.nn as nn


class PositionNet(nn.Module):
    def __init__(self, in_dim, out_dim, fourier_freqs=8):
        super().__init__()
        self.in_dim, self.out_dim, self.fourier_freqs = in_dim, out_dim, fourier_freqs
        self.position_dim = fourier_freqs * 2 * 4
        self.linears = nn.Sequential(nn.Linear(in_dim + self.position_dim, 512), nn.SiLU(),
                                     nn.Linear(512, 512), nn.SiLU(), nn.Linear(512, out_dim))
        self.null_positive_feature = nn.Parameter(torch.zeros([in_dim]))
        self.null_position_feature = nn.Parameter(torch.zeros([self.position_dim]))

    def forward(self, boxes, masks, positive_embeddings):
        raise RuntimeError("PositionNet executes inside UNetModel on the sm_100a engine (no stand-alone/CPU path)")
