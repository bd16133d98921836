"""ORACLE (test infrastructure, not product code) — plain-PyTorch restatement of the third-party
arithmetic on the hot path: diffusers==0.20.2 `UNet2DConditionModel` (requirements.txt:3 of the reference).

The reference repo contains none of this arithmetic itself; it calls it at
  trainscripts/textsliders/train_util.py:159-163 (SD1.x) and :242-247 (SDXL),
  eval-scripts/generate_images_xl.py:339-346,
and injects LoRA into its Linear/Conv2d leaves at trainscripts/textsliders/lora.py:164-218.  This file restates
the published forward of diffusers 0.20.2 (models/unet_2d_condition.py, unet_2d_blocks.py, resnet.py,
transformer_2d.py, attention.py, attention_processor.py, embeddings.py) op by op, keeping the SAME module
tree, attribute names and class names, so that
  * the reference's own lora.py (imported verbatim from /root/reference when present) discovers the same
    346 (SDXL) / 150 (SD1.x) leaves and produces the same checkpoint keys, and
  * the state-dict keys are the Hugging Face ones.

PARITY STATUS: **unpinned by the reference** — the reference ships no tests or golden tensors for this path
and diffusers is not installable here.  The restatement is pinned instead by known answers derived from the
published architecture: exact parameter counts 2 567 463 684 (SDXL-base) / 859 520 964 (SD1.x), HF key names,
LoRA leaf counts, and fp32-vs-fp64 self-consistency (tests/test_oracle.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------------
# configs (diffusers config.json of the two model families the reference trains on)
# --------------------------------------------------------------------------------------------------
@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    sample_size: int = 64
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    down_block_types: Tuple[str, ...] = ("CrossAttnDownBlock2D",) * 3 + ("DownBlock2D",)
    up_block_types: Tuple[str, ...] = ("UpBlock2D",) + ("CrossAttnUpBlock2D",) * 3
    layers_per_block: int = 2
    transformer_layers_per_block: Tuple[int, ...] = (1, 1, 1, 1)
    attention_head_dim: Tuple[int, ...] = (8, 8, 8, 8)  # diffusers quirk: this is the NUMBER of heads
    cross_attention_dim: int = 768
    use_linear_projection: bool = False
    addition_embed_type: Optional[str] = None
    addition_time_embed_dim: Optional[int] = None
    projection_class_embeddings_input_dim: Optional[int] = None
    norm_num_groups: int = 32
    norm_eps: float = 1e-5

    @staticmethod
    def sd15() -> "UNetConfig":  # CompVis/stable-diffusion-v1-4 (trainscripts/textsliders/data/config.yaml:3)
        return UNetConfig()

    @staticmethod
    def sdxl() -> "UNetConfig":  # stabilityai/stable-diffusion-xl-base-1.0 (data/config-xl.yaml:3)
        return UNetConfig(
            sample_size=128,
            block_out_channels=(320, 640, 1280),
            down_block_types=("DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"),
            up_block_types=("CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"),
            transformer_layers_per_block=(1, 2, 10),
            attention_head_dim=(5, 10, 20),
            cross_attention_dim=2048,
            use_linear_projection=True,
            addition_embed_type="text_time",
            addition_time_embed_dim=256,
            projection_class_embeddings_input_dim=2816,
        )

    @staticmethod
    def tiny_xl() -> "UNetConfig":
        """SDXL topology at 1/5 width (head dim 64 kept) — seconds on CPU; used by the parity tests."""
        return UNetConfig(
            sample_size=32,
            block_out_channels=(64, 128, 256),
            down_block_types=("DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"),
            up_block_types=("CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"),
            transformer_layers_per_block=(1, 2, 3),
            attention_head_dim=(1, 2, 4),
            cross_attention_dim=256,
            use_linear_projection=True,
            addition_embed_type="text_time",
            addition_time_embed_dim=32,
            projection_class_embeddings_input_dim=6 * 32 + 128,
        )

    @staticmethod
    def tiny_sd() -> "UNetConfig":
        """SD1.x topology at reduced width (1x1-conv projections, 4 levels, head dim 64)."""
        return UNetConfig(
            sample_size=32,
            block_out_channels=(64, 128, 256, 256),
            transformer_layers_per_block=(1, 1, 1, 1),
            attention_head_dim=(1, 2, 4, 4),
            cross_attention_dim=128,
        )


# --------------------------------------------------------------------------------------------------
# embeddings.py
# --------------------------------------------------------------------------------------------------
def get_timestep_embedding(timesteps: torch.Tensor, embedding_dim: int, flip_sin_to_cos: bool = False,
                           downscale_freq_shift: float = 1, scale: float = 1, max_period: int = 10000):
    half_dim = embedding_dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half_dim, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half_dim - downscale_freq_shift)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = scale * emb
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half_dim:], emb[:, :half_dim]], dim=-1)
    if embedding_dim % 2 == 1:
        emb = F.pad(emb, (0, 1, 0, 0))
    return emb


class Timesteps(nn.Module):
    def __init__(self, num_channels: int, flip_sin_to_cos: bool, downscale_freq_shift: float):
        super().__init__()
        self.num_channels = num_channels
        self.flip_sin_to_cos = flip_sin_to_cos
        self.downscale_freq_shift = downscale_freq_shift

    def forward(self, timesteps):
        return get_timestep_embedding(timesteps, self.num_channels, flip_sin_to_cos=self.flip_sin_to_cos,
                                      downscale_freq_shift=self.downscale_freq_shift)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels: int, time_embed_dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, sample):
        return self.linear_2(self.act(self.linear_1(sample)))


# --------------------------------------------------------------------------------------------------
# attention_processor.py / attention.py / transformer_2d.py
# --------------------------------------------------------------------------------------------------
class Attention(nn.Module):
    def __init__(self, query_dim: int, cross_attention_dim: Optional[int], heads: int, dim_head: int):
        super().__init__()
        inner_dim = dim_head * heads
        cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.to_q = nn.Linear(query_dim, inner_dim, bias=False)
        self.to_k = nn.Linear(cross_attention_dim, inner_dim, bias=False)
        self.to_v = nn.Linear(cross_attention_dim, inner_dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner_dim, query_dim), nn.Dropout(0.0)])

    def forward(self, hidden_states, encoder_hidden_states=None):
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        b, s, _ = hidden_states.shape
        q = self.to_q(hidden_states)
        k = self.to_k(ctx)
        v = self.to_v(ctx)
        h = self.heads
        d = q.shape[-1] // h
        q = q.view(b, -1, h, d).transpose(1, 2)
        k = k.view(b, -1, h, d).transpose(1, 2)
        v = v.view(b, -1, h, d).transpose(1, 2)
        # softmax(q k^T / sqrt(d)) v, written out (the reference uses xformers / SDPA; same function)
        scores = torch.matmul(q, k.transpose(-1, -2)) * self.scale
        probs = scores.softmax(dim=-1)
        o = torch.matmul(probs, v)
        o = o.transpose(1, 2).reshape(b, -1, h * d)
        o = self.to_out[0](o)
        o = self.to_out[1](o)
        return o


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, hidden_states):
        hidden_states, gate = self.proj(hidden_states).chunk(2, dim=-1)
        return hidden_states * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim: int, mult: int = 4):
        super().__init__()
        inner_dim = dim * mult
        self.net = nn.ModuleList([GEGLU(dim, inner_dim), nn.Dropout(0.0), nn.Linear(inner_dim, dim)])

    def forward(self, hidden_states):
        for module in self.net:
            hidden_states = module(hidden_states)
        return hidden_states


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, num_attention_heads: int, attention_head_dim: int, cross_attention_dim: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, num_attention_heads, attention_head_dim)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, cross_attention_dim, num_attention_heads, attention_head_dim)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, hidden_states, encoder_hidden_states):
        hidden_states = self.attn1(self.norm1(hidden_states)) + hidden_states
        hidden_states = self.attn2(self.norm2(hidden_states), encoder_hidden_states) + hidden_states
        hidden_states = self.ff(self.norm3(hidden_states)) + hidden_states
        return hidden_states


class Transformer2DModel(nn.Module):
    def __init__(self, num_attention_heads: int, attention_head_dim: int, in_channels: int, num_layers: int,
                 cross_attention_dim: int, norm_num_groups: int, use_linear_projection: bool):
        super().__init__()
        inner_dim = num_attention_heads * attention_head_dim
        self.use_linear_projection = use_linear_projection
        self.norm = nn.GroupNorm(norm_num_groups, in_channels, eps=1e-6, affine=True)
        if use_linear_projection:
            self.proj_in = nn.Linear(in_channels, inner_dim)
        else:
            self.proj_in = nn.Conv2d(in_channels, inner_dim, kernel_size=1, stride=1, padding=0)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner_dim, num_attention_heads, attention_head_dim, cross_attention_dim)
             for _ in range(num_layers)])
        if use_linear_projection:
            self.proj_out = nn.Linear(inner_dim, in_channels)
        else:
            self.proj_out = nn.Conv2d(inner_dim, in_channels, kernel_size=1, stride=1, padding=0)

    def forward(self, hidden_states, encoder_hidden_states):
        batch, _, height, width = hidden_states.shape
        residual = hidden_states
        hidden_states = self.norm(hidden_states)
        if not self.use_linear_projection:
            hidden_states = self.proj_in(hidden_states)
            inner_dim = hidden_states.shape[1]
            hidden_states = hidden_states.permute(0, 2, 3, 1).reshape(batch, height * width, inner_dim)
        else:
            inner_dim = hidden_states.shape[1]
            hidden_states = hidden_states.permute(0, 2, 3, 1).reshape(batch, height * width, inner_dim)
            hidden_states = self.proj_in(hidden_states)
        for block in self.transformer_blocks:
            hidden_states = block(hidden_states, encoder_hidden_states)
        if not self.use_linear_projection:
            hidden_states = hidden_states.reshape(batch, height, width, inner_dim).permute(0, 3, 1, 2).contiguous()
            hidden_states = self.proj_out(hidden_states)
        else:
            hidden_states = self.proj_out(hidden_states)
            hidden_states = hidden_states.reshape(batch, height, width, inner_dim).permute(0, 3, 1, 2).contiguous()
        return hidden_states + residual


# --------------------------------------------------------------------------------------------------
# resnet.py
# --------------------------------------------------------------------------------------------------
class ResnetBlock2D(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, temb_channels: int, groups: int, eps: float):
        super().__init__()
        self.norm1 = nn.GroupNorm(num_groups=groups, num_channels=in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(num_groups=groups, num_channels=out_channels, eps=eps, affine=True)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.nonlinearity = nn.SiLU()
        self.conv_shortcut = None
        if in_channels != out_channels:
            self.conv_shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=1, padding=0)

    def forward(self, input_tensor, temb):
        hidden_states = self.norm1(input_tensor)
        hidden_states = self.nonlinearity(hidden_states)
        hidden_states = self.conv1(hidden_states)
        temb = self.time_emb_proj(self.nonlinearity(temb))[:, :, None, None]
        hidden_states = hidden_states + temb
        hidden_states = self.norm2(hidden_states)
        hidden_states = self.nonlinearity(hidden_states)
        hidden_states = self.dropout(hidden_states)
        hidden_states = self.conv2(hidden_states)
        if self.conv_shortcut is not None:
            input_tensor = self.conv_shortcut(input_tensor)
        return input_tensor + hidden_states  # output_scale_factor = 1.0


class Downsample2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=1)

    def forward(self, hidden_states):
        return self.conv(hidden_states)


class Upsample2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, hidden_states):
        hidden_states = F.interpolate(hidden_states, scale_factor=2.0, mode="nearest")
        return self.conv(hidden_states)


# --------------------------------------------------------------------------------------------------
# unet_2d_blocks.py
# --------------------------------------------------------------------------------------------------
class DownBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels, num_layers, groups, eps, add_downsample):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, temb_channels, groups, eps)
            for i in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels)]) if add_downsample else None

    def forward(self, hidden_states, temb, encoder_hidden_states=None):
        output_states = ()
        for resnet in self.resnets:
            hidden_states = resnet(hidden_states, temb)
            output_states = output_states + (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            output_states = output_states + (hidden_states,)
        return hidden_states, output_states


class CrossAttnDownBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels, num_layers, transformer_layers, heads,
                 cross_attention_dim, groups, eps, add_downsample, use_linear_projection):
        super().__init__()
        resnets, attentions = [], []
        for i in range(num_layers):
            resnets.append(ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, temb_channels,
                                         groups, eps))
            attentions.append(Transformer2DModel(heads, out_channels // heads, out_channels, transformer_layers,
                                                 cross_attention_dim, groups, use_linear_projection))
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels)]) if add_downsample else None

    def forward(self, hidden_states, temb, encoder_hidden_states=None):
        output_states = ()
        for resnet, attn in zip(self.resnets, self.attentions):
            hidden_states = resnet(hidden_states, temb)
            hidden_states = attn(hidden_states, encoder_hidden_states)
            output_states = output_states + (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            output_states = output_states + (hidden_states,)
        return hidden_states, output_states


class UNetMidBlock2DCrossAttn(nn.Module):
    def __init__(self, in_channels, temb_channels, transformer_layers, heads, cross_attention_dim, groups, eps,
                 use_linear_projection):
        super().__init__()
        resnets = [ResnetBlock2D(in_channels, in_channels, temb_channels, groups, eps)]
        attentions = [Transformer2DModel(heads, in_channels // heads, in_channels, transformer_layers,
                                         cross_attention_dim, groups, use_linear_projection)]
        resnets.append(ResnetBlock2D(in_channels, in_channels, temb_channels, groups, eps))
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)

    def forward(self, hidden_states, temb, encoder_hidden_states=None):
        hidden_states = self.resnets[0](hidden_states, temb)
        for attn, resnet in zip(self.attentions, self.resnets[1:]):
            hidden_states = attn(hidden_states, encoder_hidden_states)
            hidden_states = resnet(hidden_states, temb)
        return hidden_states


class UpBlock2D(nn.Module):
    def __init__(self, in_channels, prev_output_channel, out_channels, temb_channels, num_layers, groups, eps,
                 add_upsample):
        super().__init__()
        resnets = []
        for i in range(num_layers):
            res_skip_channels = in_channels if (i == num_layers - 1) else out_channels
            resnet_in_channels = prev_output_channel if i == 0 else out_channels
            resnets.append(ResnetBlock2D(resnet_in_channels + res_skip_channels, out_channels, temb_channels,
                                         groups, eps))
        self.resnets = nn.ModuleList(resnets)
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels)]) if add_upsample else None

    def forward(self, hidden_states, res_hidden_states_tuple, temb, encoder_hidden_states=None):
        for resnet in self.resnets:
            res_hidden_states = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, res_hidden_states], dim=1)
            hidden_states = resnet(hidden_states, temb)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states)
        return hidden_states


class CrossAttnUpBlock2D(nn.Module):
    def __init__(self, in_channels, prev_output_channel, out_channels, temb_channels, num_layers,
                 transformer_layers, heads, cross_attention_dim, groups, eps, add_upsample, use_linear_projection):
        super().__init__()
        resnets, attentions = [], []
        for i in range(num_layers):
            res_skip_channels = in_channels if (i == num_layers - 1) else out_channels
            resnet_in_channels = prev_output_channel if i == 0 else out_channels
            resnets.append(ResnetBlock2D(resnet_in_channels + res_skip_channels, out_channels, temb_channels,
                                         groups, eps))
            attentions.append(Transformer2DModel(heads, out_channels // heads, out_channels, transformer_layers,
                                                 cross_attention_dim, groups, use_linear_projection))
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels)]) if add_upsample else None

    def forward(self, hidden_states, res_hidden_states_tuple, temb, encoder_hidden_states=None):
        for resnet, attn in zip(self.resnets, self.attentions):
            res_hidden_states = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, res_hidden_states], dim=1)
            hidden_states = resnet(hidden_states, temb)
            hidden_states = attn(hidden_states, encoder_hidden_states)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states)
        return hidden_states


# --------------------------------------------------------------------------------------------------
# unet_2d_condition.py
# --------------------------------------------------------------------------------------------------
class UNet2DConditionOutput(SimpleNamespace):
    pass


class UNet2DConditionModel(nn.Module):
    def __init__(self, config: UNetConfig):
        super().__init__()
        c = config
        self.config = SimpleNamespace(**c.__dict__)
        boc = c.block_out_channels
        time_embed_dim = boc[0] * 4
        self.conv_in = nn.Conv2d(c.in_channels, boc[0], kernel_size=3, padding=1)
        self.time_proj = Timesteps(boc[0], flip_sin_to_cos=True, downscale_freq_shift=0)
        self.time_embedding = TimestepEmbedding(boc[0], time_embed_dim)
        if c.addition_embed_type == "text_time":
            self.add_time_proj = Timesteps(c.addition_time_embed_dim, flip_sin_to_cos=True, downscale_freq_shift=0)
            self.add_embedding = TimestepEmbedding(c.projection_class_embeddings_input_dim, time_embed_dim)
        self.down_blocks = nn.ModuleList([])
        self.up_blocks = nn.ModuleList([])
        heads = c.attention_head_dim
        tl = c.transformer_layers_per_block
        output_channel = boc[0]
        for i, t in enumerate(c.down_block_types):
            input_channel = output_channel
            output_channel = boc[i]
            is_final = i == len(boc) - 1
            if t == "DownBlock2D":
                blk = DownBlock2D(input_channel, output_channel, time_embed_dim, c.layers_per_block,
                                  c.norm_num_groups, c.norm_eps, not is_final)
            else:
                blk = CrossAttnDownBlock2D(input_channel, output_channel, time_embed_dim, c.layers_per_block, tl[i],
                                           heads[i], c.cross_attention_dim, c.norm_num_groups, c.norm_eps,
                                           not is_final, c.use_linear_projection)
            self.down_blocks.append(blk)
        self.mid_block = UNetMidBlock2DCrossAttn(boc[-1], time_embed_dim, tl[-1], heads[-1], c.cross_attention_dim,
                                                 c.norm_num_groups, c.norm_eps, c.use_linear_projection)
        rboc = list(reversed(boc))
        rheads = list(reversed(heads))
        rtl = list(reversed(tl))
        output_channel = rboc[0]
        for i, t in enumerate(c.up_block_types):
            is_final = i == len(boc) - 1
            prev_output_channel = output_channel
            output_channel = rboc[i]
            input_channel = rboc[min(i + 1, len(boc) - 1)]
            if t == "UpBlock2D":
                blk = UpBlock2D(input_channel, prev_output_channel, output_channel, time_embed_dim,
                                c.layers_per_block + 1, c.norm_num_groups, c.norm_eps, not is_final)
            else:
                blk = CrossAttnUpBlock2D(input_channel, prev_output_channel, output_channel, time_embed_dim,
                                         c.layers_per_block + 1, rtl[i], rheads[i], c.cross_attention_dim,
                                         c.norm_num_groups, c.norm_eps, not is_final, c.use_linear_projection)
            self.up_blocks.append(blk)
        self.conv_norm_out = nn.GroupNorm(num_channels=boc[0], num_groups=c.norm_num_groups, eps=c.norm_eps)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], c.out_channels, kernel_size=3, padding=1)

    # the reference trainers call these on the diffusers model (train_lora_xl.py:78-82)
    def enable_xformers_memory_efficient_attention(self):
        return None

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    @property
    def device(self):
        return self.conv_in.weight.device

    def forward(self, sample, timestep, encoder_hidden_states, added_cond_kwargs=None,
                cross_attention_kwargs=None, return_dict: bool = True):
        timesteps = timestep
        if not torch.is_tensor(timesteps):
            # diffusers: python floats become float64 tensors, ints int64 (Euler's linspace grid is fractional)
            timesteps = torch.tensor([timesteps], dtype=torch.float64 if isinstance(timesteps, float) else torch.int64,
                                     device=sample.device)
        elif timesteps.dim() == 0:
            timesteps = timesteps[None].to(sample.device)
        timesteps = timesteps.expand(sample.shape[0])
        t_emb = self.time_proj(timesteps).to(dtype=sample.dtype)
        emb = self.time_embedding(t_emb)
        if getattr(self.config, "addition_embed_type", None) == "text_time":
            text_embeds = added_cond_kwargs["text_embeds"]
            time_ids = added_cond_kwargs["time_ids"]
            time_embeds = self.add_time_proj(time_ids.flatten())
            time_embeds = time_embeds.reshape((text_embeds.shape[0], -1))
            add_embeds = torch.concat([text_embeds, time_embeds], dim=-1).to(emb.dtype)
            emb = emb + self.add_embedding(add_embeds)

        sample = self.conv_in(sample)
        down_block_res_samples = (sample,)
        for blk in self.down_blocks:
            sample, res_samples = blk(sample, emb, encoder_hidden_states)
            down_block_res_samples += res_samples
        sample = self.mid_block(sample, emb, encoder_hidden_states)
        for blk in self.up_blocks:
            n = len(blk.resnets)
            res_samples = down_block_res_samples[-n:]
            down_block_res_samples = down_block_res_samples[:-n]
            sample = blk(sample, res_samples, emb, encoder_hidden_states)
        sample = self.conv_norm_out(sample)
        sample = self.conv_act(sample)
        sample = self.conv_out(sample)
        if not return_dict:
            return (sample,)
        return UNet2DConditionOutput(sample=sample)


def count_params(m: nn.Module) -> int:
    return sum(p.numel() for p in m.parameters())
