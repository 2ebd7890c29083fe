"""timm-free ViT / DeiT / Swin definitions with timm 0.9.2's module names.

`timm` is not installable here, and the reference's wrap_net dispatches purely on module *names* and two classes
(utils/wrap_net.py:56-64, :122-153): `qkv`, `proj`, `fc1`, `fc2`, `reduction`, `head`, `norm1`, `norm2`, `norm`,
`num_heads`, `scale`, `q_norm`, `k_norm`, `attn_drop`, `proj_drop`, `_get_rel_pos_bias`.  These stand-ins register
their children in timm's order so `named_modules()` (and therefore the calibration order)
matches.  Weights are random-init the way timm initialises them (there is no network for checkpoints).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

__all__ = ['Attention', 'Block', 'VisionTransformer', 'WindowAttention', 'SwinTransformerBlock', 'PatchMerging',
           'SwinTransformer', 'create_model', 'MODEL_ZOO']


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.drop1 = nn.Dropout(0.0)
        self.norm = nn.Identity()
        self.fc2 = nn.Linear(hidden, dim)
        self.drop2 = nn.Dropout(0.0)

    def forward(self, x):
        return self.drop2(self.fc2(self.norm(self.drop1(self.act(self.fc1(x))))))


class Attention(nn.Module):
    """timm.models.vision_transformer.Attention (0.9.2)"""

    def __init__(self, dim, num_heads=8, qkv_bias=True):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.q_norm = nn.Identity()
        self.k_norm = nn.Identity()
        self.attn_drop = nn.Dropout(0.0)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(0.0)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        q, k = self.q_norm(q), self.k_norm(k)
        attn = (q * self.scale) @ k.transpose(-2, -1)
        attn = self.attn_drop(attn.softmax(dim=-1))
        x = (attn @ v).transpose(1, 2).reshape(B, N, C)
        return self.proj_drop(self.proj(x))


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, num_heads)
        self.ls1 = nn.Identity()
        self.drop_path1 = nn.Identity()
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.ls2 = nn.Identity()
        self.drop_path2 = nn.Identity()

    def forward(self, x):
        x = x + self.drop_path1(self.ls1(self.attn(self.norm1(x))))
        return x + self.drop_path2(self.ls2(self.mlp(self.norm2(x))))


class PatchEmbed(nn.Module):
    def __init__(self, img_size, patch_size, in_chans, embed_dim, norm=False, flatten=True):
        super().__init__()
        self.img_size, self.patch_size = img_size, patch_size
        self.grid_size = img_size // patch_size
        self.num_patches = self.grid_size ** 2
        self.flatten = flatten
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = nn.LayerNorm(embed_dim) if norm else nn.Identity()

    def forward(self, x):
        x = self.proj(x)
        if self.flatten:
            x = x.flatten(2).transpose(1, 2)          # NCHW -> NLC
        else:
            x = x.permute(0, 2, 3, 1)                 # NCHW -> NHWC (Swin)
        return self.norm(x)


def _init_vit_weights(m):
    if isinstance(m, nn.Linear):
        nn.init.trunc_normal_(m.weight, std=.02)
        if m.bias is not None:
            nn.init.zeros_(m.bias)


class VisionTransformer(nn.Module):
    """ViT / DeiT (class token, global_pool='token'), timm 0.9.2 child order"""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4.0):
        super().__init__()
        self.num_classes, self.embed_dim = num_classes, embed_dim
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.randn(1, self.patch_embed.num_patches + 1, embed_dim) * .02)
        self.pos_drop = nn.Dropout(0.0)
        self.patch_drop = nn.Identity()
        self.norm_pre = nn.Identity()
        self.blocks = nn.Sequential(*[Block(embed_dim, num_heads, mlp_ratio) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self.fc_norm = nn.Identity()
        self.head_drop = nn.Dropout(0.0)
        self.head = nn.Linear(embed_dim, num_classes)
        nn.init.trunc_normal_(self.pos_embed, std=.02)
        nn.init.normal_(self.cls_token, std=1e-6)
        self.apply(_init_vit_weights)
        self.default_cfg = dict(input_size=(in_chans, img_size, img_size), mean=(0.485, 0.456, 0.406),
                                std=(0.229, 0.224, 0.225), crop_pct=0.9, interpolation='bicubic')

    def forward_features(self, x):
        x = self.patch_embed(x)
        x = torch.cat((self.cls_token.expand(x.shape[0], -1, -1), x), dim=1) + self.pos_embed
        x = self.norm_pre(self.patch_drop(self.pos_drop(x)))
        return self.norm(self.blocks(x))

    def forward_head(self, x):
        return self.head(self.head_drop(self.fc_norm(x[:, 0])))

    def forward(self, x):
        return self.forward_head(self.forward_features(x))


# ------------------------------------------------------------------------------------------------ Swin
def window_partition(x, ws):
    B, H, W, C = x.shape
    x = x.view(B, H // ws, ws, W // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, C)


def window_reverse(windows, ws, H, W):
    C = windows.shape[-1]
    x = windows.view(-1, H // ws, W // ws, ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, H, W, C)


class WindowAttention(nn.Module):
    """timm.models.swin_transformer.WindowAttention (0.9.2)"""

    def __init__(self, dim, num_heads, window_size):
        super().__init__()
        self.dim, self.num_heads = dim, num_heads
        self.window_size = (window_size, window_size)
        self.window_area = window_size * window_size
        self.scale = (dim // num_heads) ** -0.5
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * window_size - 1) ** 2, num_heads))
        coords = torch.stack(torch.meshgrid(torch.arange(window_size), torch.arange(window_size), indexing='ij'))
        cf = torch.flatten(coords, 1)
        rel = (cf[:, :, None] - cf[:, None, :]).permute(1, 2, 0).contiguous()
        rel[:, :, 0] += window_size - 1
        rel[:, :, 1] += window_size - 1
        rel[:, :, 0] *= 2 * window_size - 1
        self.register_buffer('relative_position_index', rel.sum(-1), persistent=False)
        self.qkv = nn.Linear(dim, dim * 3)
        self.attn_drop = nn.Dropout(0.0)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(0.0)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=.02)
        self.softmax = nn.Softmax(dim=-1)

    def _get_rel_pos_bias(self):
        bias = self.relative_position_bias_table[self.relative_position_index.view(-1)].view(
            self.window_area, self.window_area, -1)
        return bias.permute(2, 0, 1).contiguous().unsqueeze(0)

    def forward(self, x, mask=None):
        B_, N, C = x.shape
        qkv = self.qkv(x).reshape(B_, N, 3, self.num_heads, -1).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        attn = (q * self.scale) @ k.transpose(-2, -1) + self._get_rel_pos_bias()
        if mask is not None:
            nW = mask.shape[0]
            attn = attn.view(-1, nW, self.num_heads, N, N) + mask.unsqueeze(1).unsqueeze(0)
            attn = attn.view(-1, self.num_heads, N, N)
        attn = self.attn_drop(self.softmax(attn))
        x = (attn @ v).transpose(1, 2).reshape(B_, N, -1)
        return self.proj_drop(self.proj(x))


class SwinTransformerBlock(nn.Module):
    def __init__(self, dim, input_resolution, num_heads, window_size=7, shift_size=0, mlp_ratio=4.0):
        super().__init__()
        self.input_resolution = input_resolution
        if min(input_resolution) <= window_size:
            shift_size, window_size = 0, min(input_resolution)
        self.window_size, self.shift_size = window_size, shift_size
        self.norm1 = nn.LayerNorm(dim)
        self.attn = WindowAttention(dim, num_heads, window_size)
        self.drop_path1 = nn.Identity()
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.drop_path2 = nn.Identity()
        mask = None
        if shift_size > 0:
            H, W = input_resolution
            img = torch.zeros((1, H, W, 1))
            cnt = 0
            for h in (slice(0, -window_size), slice(-window_size, -shift_size), slice(-shift_size, None)):
                for w in (slice(0, -window_size), slice(-window_size, -shift_size), slice(-shift_size, None)):
                    img[:, h, w, :] = cnt
                    cnt += 1
            mw = window_partition(img, window_size).view(-1, window_size * window_size)
            mask = mw.unsqueeze(1) - mw.unsqueeze(2)
            mask = mask.masked_fill(mask != 0, float(-100.0)).masked_fill(mask == 0, float(0.0))
        self.register_buffer('attn_mask', mask, persistent=False)

    def _attn(self, x):
        B, H, W, C = x.shape
        if self.shift_size > 0:
            x = torch.roll(x, shifts=(-self.shift_size, -self.shift_size), dims=(1, 2))
        xw = window_partition(x, self.window_size).view(-1, self.window_size * self.window_size, C)
        aw = self.attn(xw, mask=self.attn_mask).view(-1, self.window_size, self.window_size, C)
        x = window_reverse(aw, self.window_size, H, W)
        if self.shift_size > 0:
            x = torch.roll(x, shifts=(self.shift_size, self.shift_size), dims=(1, 2))
        return x

    def forward(self, x):
        B, H, W, C = x.shape
        x = x + self.drop_path1(self._attn(self.norm1(x)))
        x = x.reshape(B, -1, C)
        x = x + self.drop_path2(self.mlp(self.norm2(x)))
        return x.reshape(B, H, W, C)


class PatchMerging(nn.Module):
    def __init__(self, dim, out_dim):
        super().__init__()
        self.norm = nn.LayerNorm(4 * dim)
        self.reduction = nn.Linear(4 * dim, out_dim, bias=False)

    def forward(self, x):
        B, H, W, C = x.shape
        x = x.reshape(B, H // 2, 2, W // 2, 2, C).permute(0, 1, 3, 4, 2, 5).flatten(3)
        return self.reduction(self.norm(x))


class SwinTransformerStage(nn.Module):
    def __init__(self, dim, out_dim, input_resolution, depth, downsample, num_heads, window_size, mlp_ratio):
        super().__init__()
        self.downsample = PatchMerging(dim, out_dim) if downsample else nn.Identity()
        res = tuple(i // 2 for i in input_resolution) if downsample else input_resolution
        self.blocks = nn.Sequential(*[
            SwinTransformerBlock(out_dim, res, num_heads, window_size, 0 if i % 2 == 0 else window_size // 2, mlp_ratio)
            for i in range(depth)])

    def forward(self, x):
        return self.blocks(self.downsample(x))


class ClassifierHead(nn.Module):
    def __init__(self, in_features, num_classes):
        super().__init__()
        self.global_pool = nn.Identity()
        self.drop = nn.Dropout(0.0)
        self.fc = nn.Linear(in_features, num_classes)
        self.flatten = nn.Identity()

    def forward(self, x):
        return self.flatten(self.fc(self.drop(x.mean(dim=(1, 2)))))


class SwinTransformer(nn.Module):
    def __init__(self, img_size=224, patch_size=4, in_chans=3, num_classes=1000, embed_dim=96, depths=(2, 2, 6, 2),
                 num_heads=(3, 6, 12, 24), window_size=7, mlp_ratio=4.0):
        super().__init__()
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim, norm=True, flatten=False)
        grid = (img_size // patch_size, img_size // patch_size)
        dims = [embed_dim * 2 ** i for i in range(len(depths))]
        layers, in_dim, res = [], embed_dim, grid
        for i, d in enumerate(depths):
            layers.append(SwinTransformerStage(in_dim, dims[i], res, d, i > 0, num_heads[i], window_size, mlp_ratio))
            if i > 0:
                res = (res[0] // 2, res[1] // 2)
            in_dim = dims[i]
        self.layers = nn.Sequential(*layers)
        self.norm = nn.LayerNorm(dims[-1])
        self.head = ClassifierHead(dims[-1], num_classes)
        self.apply(_init_vit_weights)
        self.default_cfg = dict(input_size=(in_chans, img_size, img_size), mean=(0.485, 0.456, 0.406),
                                std=(0.229, 0.224, 0.225), crop_pct=0.9, interpolation='bicubic')

    def forward(self, x):
        return self.head(self.norm(self.layers(self.patch_embed(x))))


MODEL_ZOO = {
    'vit_tiny_patch16_224': lambda: VisionTransformer(embed_dim=192, depth=12, num_heads=3),
    'vit_small_patch16_224': lambda: VisionTransformer(embed_dim=384, depth=12, num_heads=6),
    'vit_base_patch16_224': lambda: VisionTransformer(embed_dim=768, depth=12, num_heads=12),
    'vit_large_patch16_224': lambda: VisionTransformer(embed_dim=1024, depth=24, num_heads=16),
    'deit_tiny_patch16_224': lambda: VisionTransformer(embed_dim=192, depth=12, num_heads=3),
    'deit_small_patch16_224': lambda: VisionTransformer(embed_dim=384, depth=12, num_heads=6),
    'deit_base_patch16_224': lambda: VisionTransformer(embed_dim=768, depth=12, num_heads=12),
    'swin_tiny_patch4_window7_224': lambda: SwinTransformer(embed_dim=96, depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24)),
    'swin_small_patch4_window7_224': lambda: SwinTransformer(embed_dim=96, depths=(2, 2, 18, 2), num_heads=(3, 6, 12, 24)),
    'swin_base_patch4_window7_224': lambda: SwinTransformer(embed_dim=128, depths=(2, 2, 18, 2), num_heads=(4, 8, 16, 32)),
    'swin_base_patch4_window12_384': lambda: SwinTransformer(img_size=384, window_size=12, embed_dim=128,
                                                             depths=(2, 2, 18, 2), num_heads=(4, 8, 16, 32)),
    # one-block DeiT-Tiny: the ncu launch-list target (profiles/)
    'deit_tiny_depth1_patch16_224': lambda: VisionTransformer(embed_dim=192, depth=1, num_heads=3),
    'deit_small_depth2_patch16_224': lambda: VisionTransformer(embed_dim=384, depth=2, num_heads=6),
    # patch embedding + block 0 + head of DeiT-Small: the free-running parity slice of BASELINE config 2
    'deit_small_depth1_patch16_224': lambda: VisionTransformer(embed_dim=384, depth=1, num_heads=6),
    # tiny configurations for tests
    'vit_test_patch8_32': lambda: VisionTransformer(img_size=32, patch_size=8, num_classes=10, embed_dim=32, depth=2,
                                                    num_heads=2),
    'swin_test_patch2_window4_32': lambda: SwinTransformer(img_size=32, patch_size=2, num_classes=10, embed_dim=16,
                                                           depths=(2, 2), num_heads=(2, 4), window_size=4),
}


def create_model(name, pretrained=False, checkpoint_path=None, **kwargs):
    """timm.create_model stand-in: random-init weights (same seed => same weights), optional state_dict file."""
    if name not in MODEL_ZOO:
        raise ValueError(f'unknown model {name}')
    model = MODEL_ZOO[name]()
    if checkpoint_path:
        model.load_state_dict(torch.load(checkpoint_path, map_location='cpu'))
    return model
