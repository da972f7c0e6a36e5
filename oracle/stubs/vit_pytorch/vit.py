"""TEST INFRASTRUCTURE (oracle) — restatement of the parts of `vit-pytorch==1.6.4`
(pinned at /root/reference/requirements.txt:160; the package itself is NOT in the reference tree
and not installable offline) that the reference's hot path calls:
`pair` and `Transformer` (/root/reference/models/pretrain_models.py:1,27,113,784).

PARITY UNPINNED for this file: it is written from the package's published algorithm (pre-norm
attention + feed-forward residual blocks, final LayerNorm), not checked against the wheel.  The
reference tree pins only its call signature `Transformer(dim, depth, heads, dim_head, mlp_dim,
dropout)` (/root/reference/dino_test.py:8) and, through saved checkpoints, the parameter names.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this.  It exists so that /root/reference/models/pretrain_models.py can be imported
unmodified for cross-checks.
"""
import torch
from torch import nn


def pair(t):
    return t if isinstance(t, tuple) else (t, t)


class FeedForward(nn.Module):
    # LayerNorm -> Linear -> GELU(erf) -> Dropout -> Linear -> Dropout ; indices 0,1,2,3,4,5
    def __init__(self, dim, hidden_dim, dropout=0.0):
        super().__init__()
        self.net = nn.Sequential(
            nn.LayerNorm(dim),
            nn.Linear(dim, hidden_dim),
            nn.GELU(),
            nn.Dropout(dropout),
            nn.Linear(hidden_dim, dim),
            nn.Dropout(dropout),
        )

    def forward(self, x):
        return self.net(x)


class Attention(nn.Module):
    def __init__(self, dim, heads=8, dim_head=64, dropout=0.0):
        super().__init__()
        inner = dim_head * heads
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.norm = nn.LayerNorm(dim)
        self.attend = nn.Softmax(dim=-1)
        self.dropout = nn.Dropout(dropout)
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        project_out = not (heads == 1 and dim_head == dim)
        self.to_out = (
            nn.Sequential(nn.Linear(inner, dim), nn.Dropout(dropout)) if project_out else nn.Identity()
        )

    def forward(self, x):
        b, n, _ = x.shape
        x = self.norm(x)
        q, k, v = self.to_qkv(x).chunk(3, dim=-1)
        # 'b n (h d) -> b h n d'
        q, k, v = (t.reshape(b, n, self.heads, -1).transpose(1, 2) for t in (q, k, v))
        dots = torch.matmul(q, k.transpose(-1, -2)) * self.scale
        attn = self.dropout(self.attend(dots))
        out = torch.matmul(attn, v)
        out = out.transpose(1, 2).reshape(b, n, -1)  # 'b h n d -> b n (h d)'
        return self.to_out(out)


class Transformer(nn.Module):
    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout=0.0):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(
                nn.ModuleList(
                    [
                        Attention(dim, heads=heads, dim_head=dim_head, dropout=dropout),
                        FeedForward(dim, mlp_dim, dropout=dropout),
                    ]
                )
            )

    def forward(self, x):
        for attn, ff in self.layers:
            x = attn(x) + x
            x = ff(x) + x
        return self.norm(x)
