"""TEST INFRASTRUCTURE (oracle). ``diffusers.models.attention_processor.Attention`` attribute contract
restated from diffusers==0.24.0 (not under /root/reference; parity unpinned): the reference reads
``heads, to_q, to_k, to_v, to_out, residual_connection, rescale_output_factor``
(/root/reference/distrifuser/modules/pp/attn.py:16-38,93-100)."""
import torch
from torch import nn
from torch.nn import functional as F


class Attention(nn.Module):
    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, bias=False, out_bias=True):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.inner_dim = inner
        self.scale = dim_head ** -0.5
        self.residual_connection = False
        self.rescale_output_factor = 1.0
        kv_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(kv_dim, inner, bias=bias)
        self.to_v = nn.Linear(kv_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=out_bias), nn.Dropout(0.0)])

    def forward(self, hidden_states, encoder_hidden_states=None, **kwargs):
        # AttnProcessor2_0 of diffusers 0.24.0 (no mask, no group norm, no added kv)
        b = hidden_states.shape[0]
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        q, k, v = self.to_q(hidden_states), self.to_k(ctx), self.to_v(ctx)
        d = self.inner_dim // self.heads
        q = q.view(b, -1, self.heads, d).transpose(1, 2)
        k = k.view(b, -1, self.heads, d).transpose(1, 2)
        v = v.view(b, -1, self.heads, d).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(b, -1, self.inner_dim).to(q.dtype)
        o = self.to_out[1](self.to_out[0](o))
        if self.residual_connection:
            o = o + hidden_states
        return o / self.rescale_output_factor
