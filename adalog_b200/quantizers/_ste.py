"""Straight-through estimators (reference quantizers/_ste.py:5-14); only live in training_mode (BRECQ)."""
import torch


def round_ste(x: torch.Tensor):
    return (x.round() - x).detach() + x


def floor_ste(x: torch.Tensor):
    return (x.floor() - x).detach() + x


def ceil_ste(x: torch.Tensor):
    return (x.ceil() - x).detach() + x
