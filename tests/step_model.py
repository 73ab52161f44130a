"""A tiny stand-in for the reference U-Net used by the train-step tests and their golden generator: two 3x3 convolutions
and the reference's output-activation switch (src/unet.py:156-167, :208-214: `activation_name` selects sigmoid or
(tanh+1)/2 inside forward; any other value returns the logits)."""
import torch
import torch.nn as nn


class TinySegNet(nn.Module):
    def __init__(self, hidden: int = 4, activation: str = "sigmoid"):
        super().__init__()
        self.c1 = nn.Conv2d(1, hidden, 3, padding=1)
        self.c2 = nn.Conv2d(hidden, 1, 3, padding=1)
        self.activation_name = activation

    def forward(self, x):
        z = self.c2(torch.relu(self.c1(x)))
        if self.activation_name == "sigmoid":
            return torch.sigmoid(z)
        if self.activation_name == "tanh":
            return (torch.tanh(z) + 1.0) / 2.0
        return z


def make_batches(n_batches: int, B: int, H: int, W: int, seed: int):
    """Deterministic (images, masks) batches: smooth images, masks correlated with them."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n_batches):
        lo = torch.randn(B, 1, max(H // 8, 2), max(W // 8, 2), generator=g)
        sm = torch.nn.functional.interpolate(lo, size=(H, W), mode="bicubic", align_corners=False)
        images = (sm + 0.2 * torch.randn(B, 1, H, W, generator=g)).contiguous()
        masks = (sm > 0.2).float().contiguous()
        out.append((images, masks))
    return out
