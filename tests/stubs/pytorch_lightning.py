"""Test stand-in for `pytorch_lightning` (tools/aesthetic.py subclasses LightningModule)."""
import torch

LightningModule = torch.nn.Module
