"""CvT-13-shaped stub for the reference's `architectures` package (test_diml_cvt.py:42,82), which
cannot be imported here (timm, pretrainedmodels, missing cvt_cross, hard-coded checkpoint paths).
select(arch, opt) returns a module with exactly what evaluate() touches
(evaluation/eval_cvt_diml.py:205-206,251-279 and get_attention_rollout :111-125):

    .pars                         the argparse namespace
    __call__(img) -> (emb [B, embed_dim], (enc_out [B, 384], tokens [B, 196, 384]))
    .model.head                   Linear(384 -> embed_dim)
    .model.both_forward(img)      fills stage{0,1,2}.blocks[i]._probs[0] = attention [B, heads, T, T]
"""
import torch
import torch.nn as nn


class _Block(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.heads = heads
        self.qk = nn.Linear(dim, 2 * dim)
        self._probs = [None]

    def forward(self, x):
        b, t, d = x.shape
        q, k = self.qk(x).reshape(b, t, 2, self.heads, d // self.heads).permute(2, 0, 3, 1, 4)
        att = torch.softmax(q @ k.transpose(-1, -2) / (d // self.heads) ** 0.5, dim=-1)
        self._probs = [att]
        return x


class _Stage(nn.Module):
    def __init__(self, blocks):
        super().__init__()
        self.blocks = nn.ModuleList(blocks)


class _Backbone(nn.Module):
    def __init__(self, embed_dim):
        super().__init__()
        self.patch = nn.Conv2d(3, 384, kernel_size=16, stride=16)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, 384))
        self.stage0 = _Stage([])
        self.stage1 = _Stage([])
        self.stage2 = _Stage([_Block(384, 6), _Block(384, 6)])
        self.head = nn.Linear(384, embed_dim)

    def tokens(self, img):
        return self.patch(img).flatten(2).transpose(1, 2)          # [B, 196, 384]

    def both_forward(self, img):
        t = self.tokens(img)
        x = torch.cat([self.cls_token.expand(t.size(0), -1, -1), t], dim=1)
        for blk in self.stage2.blocks:
            x = blk(x)
        return t, x[:, :1]


class Network(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.pars = opt
        self.name = getattr(opt, 'arch', 'cvt_13_normalize')
        self.model = _Backbone(getattr(opt, 'embed_dim', 128))

    def forward(self, img):
        tokens = self.model.tokens(img)
        enc = tokens.mean(dim=1)
        emb = self.model.head(enc)
        if 'normalize' in self.name:
            emb = torch.nn.functional.normalize(emb, dim=-1)
        return emb, (enc, tokens)


def select(arch, opt):
    torch.manual_seed(getattr(opt, 'seed', 0))
    return Network(opt)
