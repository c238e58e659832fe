"""Host-side mirror of the reference ``VQVAE`` around the B200 quantizer (boundary rows of SURVEY.md section 8a:
a13 quantizer call site, a14 vq-loss term, a15 stats / sample / codebook init).

Scope.  The encoder, tokenizer and decoder are stock PyTorch modules (not the hot path); they are laid out
with exactly the reference's module and buffer names (``models/vq_vae.py:455-553``) so that a reference
checkpoint loads with ``strict=True``.  ``forward`` / ``encode`` / ``decode`` / ``sample`` /
``init_codebook_from_centroids`` / ``_compute_stats`` follow the reference call contract.  ``loss_function``
implements the terms that touch the quantizer or the plain reconstruction (raw xyz MSE, secondary-structure
CE, ``VQ_Loss``, total-variation of the SS probabilities); the reference's Kabsch / bond / Frenet / PDM
geometry regularisers (``models/vq_vae.py:903-1095,1138-1290,1311-1330``) are CPU-friendly auxiliary
PyTorch code outside this package's scope: asking for them raises, and the supported way to train with
them is the reference's own ``VQVAE`` with this quantizer bound in by ``pytorch_vae_b200.install()``.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .quantizer import VectorQuantizerEMA

Tensor = torch.Tensor

_GEOMETRY_WEIGHTS = ("bond_length_weight", "bond_angle_weight", "xyz_tv_lambda", "dir_weight", "dih_weight",
                     "pdm_weight", "win_kabsch_weight", "kappa_weight", "tau_weight", "lr_pdm_weight")


class LatentTokenizer(nn.Module):
    """Learned-query cross-attention L -> N tokens; names as in models/vq_vae.py:288-322."""

    def __init__(self, d_model: int, n_tokens: int = 32, n_heads: int = 8, n_layers: int = 2, dropout: float = 0.1):
        super().__init__()
        self.n_tokens, self.d = int(n_tokens), int(d_model)
        self.queries = nn.Parameter(torch.randn(self.n_tokens, self.d) * 0.02)
        self.drop = nn.Dropout(dropout)
        self.layers = nn.ModuleList(
            nn.ModuleDict({
                "ln_q": nn.LayerNorm(self.d),
                "ln_kv": nn.LayerNorm(self.d),
                "attn": nn.MultiheadAttention(self.d, int(n_heads), batch_first=True, dropout=dropout),
                "ln_o": nn.LayerNorm(self.d),
                "ffn": nn.Sequential(nn.Linear(self.d, 4 * self.d), nn.GELU(), nn.Linear(4 * self.d, self.d)),
                "ffn_drop": nn.Dropout(dropout),
            }) for _ in range(int(n_layers)))

    def forward(self, x: Tensor, key_padding_mask: Optional[Tensor] = None) -> Tensor:
        q = self.queries.unsqueeze(0).expand(x.size(0), -1, -1)
        for blk in self.layers:
            kv = blk["ln_kv"](x)
            att, _ = blk["attn"](blk["ln_q"](q), kv, kv, key_padding_mask=key_padding_mask, need_weights=False)
            q = q + self.drop(att)
            q = q + blk["ffn_drop"](blk["ffn"](blk["ln_o"](q)))
        return q


class VQVAE(nn.Module):
    """Curve VQ-VAE with the B200 quantizer; constructor keywords as ``models/vq_vae.py:366-408``
    (unknown keys, e.g. ``name``, are swallowed like the reference's ``**kwargs``)."""

    def __init__(self, input_dim: int = 6, hidden_dim: int = 512, num_layers: int = 4, num_heads: int = 8,
                 max_seq_len: int = 350, codebook_size: int = 512, code_dim: int = 128, beta: float = 0.25,
                 use_vq: bool = True, residual_vq: bool = False, num_quantizers: int = 1,
                 label_smoothing: float = 0.0, ss_tv_lambda: float = 0.0, usage_entropy_lambda: float = 0.0,
                 xyz_align_alpha: float = 0.7, codebook_init_path: Optional[str] = None,
                 ema_decay_start: float = 0.98, ema_decay_end: float = 0.98, ema_decay_warm_steps: int = 0,
                 soft_vq_use: bool = False, soft_vq_tau_start: float = 2.0, soft_vq_tau_end: float = 0.5,
                 soft_vq_tau_warm_steps: int = 0, soft_vq_alpha_warm_steps: int = 0,
                 latent_tokens: int = 32, tokenizer_heads: int = 8,
                 tokenizer_layers: int = 2, tokenizer_dropout: float = 0.1, latent_sigmoid: bool = False,
                 latent_sigmoid_ae_only: bool = True, reinit_dead_codes: bool = True, reinit_prob: float = 1.0,
                 dead_usage_threshold: int = 0, ema_update_freeze_steps: int = 0, print_init: bool = True,
                 search_mode: str = "fp32", rigid_aug_prob: float = 0.0, max_noise_std: float = 0.0,
                 noise_warmup_steps: int = 0, **kwargs):
        super().__init__()
        # input augmentations of the reference's forward (models/vq_vae.py:775-792) are not on the hot path and are
        # not mirrored: refuse them loudly instead of silently training without them
        if float(rigid_aug_prob) != 0.0 or float(max_noise_std) != 0.0:
            raise NotImplementedError("rigid_aug_prob / max_noise_std != 0: input augmentation is outside the hot path; "
                                      "train with the reference VQVAE and pytorch_vae_b200.install()")
        self.rigid_aug_prob, self.max_noise_std = 0.0, 0.0
        self.noise_warmup_steps = int(noise_warmup_steps)
        self.soft_vq_use = bool(soft_vq_use)                # models/vq_vae.py:436-440
        self.soft_vq_tau_start, self.soft_vq_tau_end = float(soft_vq_tau_start), float(soft_vq_tau_end)
        self.soft_vq_tau_warm_steps = int(soft_vq_tau_warm_steps)
        self.soft_vq_alpha_warm_steps = int(soft_vq_alpha_warm_steps)
        self.input_dim, self.hidden_dim = int(input_dim), int(hidden_dim)
        self.code_dim, self.max_seq_len = int(code_dim), int(max_seq_len)
        self.use_vq = bool(use_vq)
        self._beta = float(beta)
        self.num_quantizers = int(num_quantizers)
        self.residual_vq = self.use_vq and self.num_quantizers > 1          # derived, the flag itself is ignored
        self.label_smoothing = float(label_smoothing)
        self.ss_tv_lambda = float(ss_tv_lambda)
        self.usage_entropy_lambda = float(usage_entropy_lambda)
        self.xyz_align_alpha = float(xyz_align_alpha)
        self.ema_decay_start, self.ema_decay_end = float(ema_decay_start), float(ema_decay_end)
        self.ema_decay_warm_steps = int(ema_decay_warm_steps)
        self.ema_update_freeze_steps = int(ema_update_freeze_steps)
        self.codebook_init_path = codebook_init_path
        self.latent_n_tokens = int(latent_tokens)
        self.latent_sigmoid, self.latent_sigmoid_ae_only = bool(latent_sigmoid), bool(latent_sigmoid_ae_only)
        self.training_steps = 0
        self._curr_epoch = 0
        self._ema_decay_override = None

        H = self.hidden_dim

        def enc_stack(layers):
            layer = nn.TransformerEncoderLayer(d_model=H, nhead=num_heads, batch_first=True, dropout=0.1, norm_first=True)
            return nn.TransformerEncoder(layer, num_layers=layers)

        self.input_proj = nn.Linear(3, H)
        self.ss_input_proj = nn.Linear(3, H)
        self.inp_dropout = nn.Dropout(p=0.1)
        self.encoder = enc_stack(num_layers)
        self.enc_ln = nn.LayerNorm(H)
        self.to_code = nn.Linear(H, self.code_dim)
        self.ln_geo = nn.LayerNorm(H)
        self.ln_ss = nn.LayerNorm(H)
        self.ss_encoder = enc_stack(2)

        pos = torch.arange(self.max_seq_len, dtype=torch.float32).unsqueeze(1)
        freq = torch.exp(torch.arange(0, H, 2).float() * (-math.log(10000.0) / H))
        pe = torch.zeros(self.max_seq_len, H)
        pe[:, 0::2], pe[:, 1::2] = torch.sin(pos * freq), torch.cos(pos * freq)
        self.register_buffer("pos_enc", pe.unsqueeze(0))

        self.tokenizer = LatentTokenizer(H, self.latent_n_tokens, int(tokenizer_heads), int(tokenizer_layers),
                                         float(tokenizer_dropout))
        self.fuse_mlp = nn.Sequential(nn.Linear(2 * H, H), nn.GELU(), nn.Linear(H, H), nn.LayerNorm(H))

        if self.use_vq:
            self.quantizer = VectorQuantizerEMA(
                num_embeddings=codebook_size, embedding_dim=self.code_dim, beta=beta, decay=0.98, eps=1e-5,
                reinit_dead_codes=reinit_dead_codes, reinit_prob=reinit_prob,
                dead_usage_threshold=dead_usage_threshold, print_init=print_init,
                num_quantizers=self.num_quantizers, search_mode=search_mode)
            self.quantizer.beta = self._beta
        else:
            self.quantizer = None

        self.from_code = nn.Linear(self.code_dim, H)
        self.mem_ln = nn.LayerNorm(H)
        dec_layer = nn.TransformerDecoderLayer(d_model=H, nhead=num_heads, batch_first=True, dropout=0.1, norm_first=True)
        self.decoder = nn.TransformerDecoder(dec_layer, num_layers=num_layers)
        self.query_embed = nn.Embedding(self.max_seq_len, H)
        nn.init.normal_(self.query_embed.weight, std=0.02)
        self.head_xyz = nn.Linear(H, 3)
        self.head_ss = nn.Linear(H, 3)

        if self.use_vq and self.codebook_init_path:
            self.init_codebook_from_centroids(torch.from_numpy(np.load(self.codebook_init_path).astype(np.float32)))

    # ---------------------------------------------------------------- small surface used by the harness
    @property
    def beta(self):
        return self._beta

    @beta.setter
    def beta(self, value):                                   # models/vq_vae.py:559-563
        self._beta = float(value)
        if self.quantizer is not None:
            self.quantizer.beta = float(value)

    def set_epoch_context(self, epoch: int, steps_per_epoch: int = 1):
        self._curr_epoch = int(epoch)

    @torch.no_grad()
    def init_codebook_from_centroids(self, centroids: Tensor):
        """[K_total, D] or [L, K_per, D] centroids -> embedding = ema_embedding = C, ema_cluster_size = 1
        (models/vq_vae.py:577-613)."""
        q = self.quantizer
        if q is None:
            raise ValueError("Quantizer is not initialized.")
        if centroids.dim() == 3:
            L, K_per, D = centroids.shape
            if D != self.code_dim:
                raise ValueError(f"Centroid D mismatch: expected {self.code_dim}, got {D}")
            if L * K_per != q.K:
                raise ValueError(f"Centroid K mismatch: expected {q.K}, got {L * K_per}")
            flat = centroids.reshape(-1, D)
        elif centroids.dim() == 2:
            if tuple(centroids.shape) != (q.K, self.code_dim):
                raise ValueError(f"Centroid shape mismatch: expected {(q.K, self.code_dim)}, got {tuple(centroids.shape)}")
            flat = centroids
        else:
            raise ValueError(f"Unsupported centroid shape: {tuple(centroids.shape)}")
        flat = flat.to(device=q.embedding.device, dtype=q.embedding.dtype)
        q.embedding.copy_(flat)                              # bumps the version counter: the kernel cache refreshes
        q.ema_embedding.copy_(flat)
        q.ema_cluster_size.fill_(1.0)

    def _compute_stats(self, indices: Tensor, device) -> Tuple[Tensor, Tensor]:
        """Perplexity / dead ratio of an index tensor through the fused statistics kernel (models/vq_vae.py:627-637)."""
        if self.quantizer is None:
            return torch.tensor(0.0, device=device), torch.tensor(0.0, device=device)
        with torch.no_grad():
            hist = torch.bincount(indices.reshape(-1), minlength=self.quantizer.K).to(torch.int32)
            out = torch.empty(3, dtype=torch.float32, device=hist.device)
            ops.stats_finalize(hist, 0.0, None, 0.0, None, None, out)
        return out[0], out[1]

    # ---------------------------------------------------------------- encoder / decoder (stock PyTorch)
    def encode(self, x: Tensor, mask: Optional[Tensor] = None):
        L = x.size(1)
        pad = (~mask) if mask is not None else None
        pos = self.pos_enc[:, :L, :]
        geo = self.encoder(self.inp_dropout(self.input_proj(x[..., :3])) + pos, src_key_padding_mask=pad)
        h_enc_geo = self.enc_ln(geo)
        h_enc_ss = self.ss_encoder(self.ss_input_proj(x[..., 3:]) + pos, src_key_padding_mask=pad)
        fused = self.fuse_mlp(torch.cat([self.ln_geo(h_enc_geo), self.ln_ss(h_enc_ss)], dim=-1))
        return fused, h_enc_geo, h_enc_ss

    def _tokenize_to_codes(self, h_tokens: Tensor, mask: Optional[Tensor]) -> Tensor:
        z = self.to_code(self.tokenizer(h_tokens, key_padding_mask=(~mask) if mask is not None else None))
        if self.latent_sigmoid and ((not self.latent_sigmoid_ae_only) or (not self.use_vq)):
            z = torch.sigmoid(z)
        return z

    def decode(self, z_for_decode: Tensor, mask: Optional[Tensor] = None) -> Tensor:
        return self._decode_memory(self.mem_ln(self.from_code(z_for_decode)), mask)

    def _decode_memory(self, memory: Tensor, mask: Optional[Tensor]) -> Tensor:
        B = memory.size(0)
        L = mask.size(1) if mask is not None else self.max_seq_len
        q = self.query_embed.weight[:L].unsqueeze(0).expand(B, L, -1) + self.pos_enc[:, :L, :]
        h = self.decoder(tgt=q, memory=memory, tgt_key_padding_mask=(~mask) if mask is not None else None)
        return torch.cat([self.head_xyz(h), self.head_ss(h)], dim=-1)

    # ---- indices -> decoder memory without z_q and without a GEMM over the tokens (SURVEY.md section 8f rank 1:
    # "from_code + mem_ln after K2"; the inference path of scripts/decode_with_vqvae.py:110-130 + models/vq_vae.py:749)
    @torch.no_grad()
    def projected_codebook(self) -> Tensor:
        """P = embedding @ from_code.weight^T  [K_total, hidden]: from_code(sum_q E[i_q]) = sum_q P[i_q] + bias.
        Rebuilt only when the codebook or the weight has changed (tensor version counters)."""
        q = self.quantizer
        key = (q.embedding._version, self.from_code.weight._version, q.embedding.data_ptr(), self.from_code.weight.data_ptr())
        cached = getattr(self, "_proj_cache", None)
        if cached is None or cached[0] != key:
            P = (q.embedding.double() @ self.from_code.weight.double().t()).float().contiguous()   # once per version
            object.__setattr__(self, "_proj_cache", (key, P))
            cached = self._proj_cache
        return cached[1]

    @torch.no_grad()
    def memory_from_indices(self, indices: Tensor) -> Tensor:
        """Token-major global ids [B, M*Q] / [B, M, Q] -> memory [B, M, hidden] = mem_ln(from_code(z_q)) in ONE
        gather + LayerNorm kernel (inference only: no gradient)."""
        q = self.quantizer
        Q = int(q.num_quantizers)
        B = indices.size(0)
        idx = indices.reshape(B, -1)
        if idx.size(1) % Q != 0:
            raise ValueError(f"index row length {idx.size(1)} is not divisible by num_quantizers={Q}")
        mem = ops.indices_to_memory(idx, self.projected_codebook(), Q, self.from_code.bias,
                                    self.mem_ln.weight if self.mem_ln.elementwise_affine else None,
                                    self.mem_ln.bias if self.mem_ln.elementwise_affine else None, self.mem_ln.eps)
        return mem.view(B, idx.size(1) // Q, -1)

    @torch.no_grad()
    def decode_indices(self, indices: Tensor, mask: Optional[Tensor] = None) -> Tensor:
        """decode(indices_to_latent(indices)) without forming the latent."""
        return self._decode_memory(self.memory_from_indices(indices), mask)

    # ---------------------------------------------------------------- forward: the quantizer call site
    def forward(self, x: Tensor, mask: Optional[Tensor] = None, **kwargs) -> List[Tensor]:
        target = x.clone()
        if self.quantizer is not None:                       # EMA decay schedule, models/vq_vae.py:794-802
            if self._ema_decay_override is not None:
                self.quantizer.decay = float(self._ema_decay_override)
            else:
                w = self.ema_decay_warm_steps
                t = 1.0 if w <= 0 else min(1.0, max(0.0, self.training_steps) / float(w))
                self.quantizer.decay = float((1.0 - t) * self.ema_decay_start + t * self.ema_decay_end) if w > 0 \
                    else float(self.ema_decay_end)
        h_fuse, _, _ = self.encode(x, mask=mask)
        if self.training:
            self.training_steps += 1
        z_e = self._tokenize_to_codes(h_fuse, mask)

        do_ema = False
        if not self.use_vq or self.quantizer is None:
            z_dec, z_q_raw = z_e, z_e
            indices = torch.zeros(z_e.size(0), z_e.size(1), dtype=torch.long, device=z_e.device)
            ppl = dead = torch.tensor(0.0, device=x.device)
        elif self.soft_vq_use and self.training and not self.residual_vq:
            # soft VQ, single-level only (models/vq_vae.py:828-861): decode a blend of the softmax-weighted code
            # mixture and the hard code; the EMA update and the statistics use the hard assignment
            do_ema = self.training_steps >= self.ema_update_freeze_steps
            w = self.soft_vq_tau_warm_steps                                   # _interp_linear(:621-625)
            t = 1.0 if w <= 0 else min(1.0, max(0.0, self.training_steps) / float(w))
            tau = self.soft_vq_tau_end if w <= 0 else (1.0 - t) * self.soft_vq_tau_start + t * self.soft_vq_tau_end
            aw = self.soft_vq_alpha_warm_steps                                # _linear_schedule(:615-619)
            alpha = 1.0 if aw <= 0 else min(1.0, float(self.training_steps) / float(aw))
            z_soft, z_q_raw, indices, stats = self.quantizer.soft_forward(z_e, tau, do_ema_update=do_ema)
            z_mix = (1 - alpha) * z_soft + alpha * z_q_raw
            z_dec = z_e + (z_mix - z_e).detach()
            ppl, dead = stats[0], stats[1]
        else:
            do_ema = self.training and self.training_steps >= self.ema_update_freeze_steps
            z_dec, z_q_raw, indices, stats = self.quantizer(z_e, do_ema_update=do_ema, allow_reinit=do_ema, mask=None)
            ppl, dead = stats[0], stats[1]
        # dead-code re-init cadence, models/vq_vae.py:874-891: AFTER the soft / hard branch, for both of them
        # (every 500 steps once training_steps >= max(freeze, 800)), gated on the branch's do_ema_update
        if self.use_vq and self.quantizer is not None and self.training and do_ema and \
                self.training_steps % 500 == 0 and self.training_steps >= max(self.ema_update_freeze_steps, 800):
            usage = torch.bincount(indices.reshape(-1), minlength=self.quantizer.K).float()
            self.quantizer._maybe_reinit_dead_codes(z_e.detach().reshape(-1, z_e.size(-1)), usage)

        recons = self.decode(z_dec, mask=mask)
        return [recons, target, (z_q_raw, z_e, indices, ppl, dead), mask]

    # ---------------------------------------------------------------- loss: VQ term + plain reconstruction
    def loss_function(self, *args, **kwargs) -> dict:
        recons, target, vq_pack = args[0], args[1], args[2]
        mask = args[3] if len(args) > 3 else None
        zq_raw, ze_raw, _indices, ppl, dead = vq_pack
        for name in _GEOMETRY_WEIGHTS:
            if float(kwargs.get(name, 0.0)) != 0.0:
                raise NotImplementedError(f"{name} != 0: the geometry regularisers are outside the hot path; train with "
                                          "the reference VQVAE and pytorch_vae_b200.install()")
        if self.xyz_align_alpha != 0.0:
            raise NotImplementedError("Kabsch-aligned xyz loss: use the reference VQVAE with pytorch_vae_b200.install()")
        ss_weight, rmsd_weight = float(kwargs.get("ss_weight", 1.0)), float(kwargs.get("rmsd_weight", 1.0))
        dev = recons.device
        zero = torch.tensor(0.0, device=dev)
        re_xyz, logits = recons[..., :3], recons[..., 3:]
        gt_xyz, labels = target[..., :3], target[..., 3:].argmax(-1)

        d2 = (re_xyz - gt_xyz).pow(2).sum(-1)
        if mask is None:
            per_sample = d2.mean(1)
        else:
            m = mask.float()
            per_sample = (d2 * m).sum(1) / m.sum(1).clamp_min(1.0)
        loss_xyz = per_sample.mean()
        rmsd = torch.sqrt(per_sample.detach().clamp_min(1e-12)).mean()

        logp = F.log_softmax(logits, dim=-1)
        logp_y = logp.gather(-1, labels.unsqueeze(-1)).squeeze(-1)
        if self.label_smoothing > 0.0:
            # KL(t || p) with t = 1-eps on the label and eps/(C-1) elsewhere (models/vq_vae.py:920-931), closed form
            C = logits.size(-1)
            on, off = 1.0 - self.label_smoothing, self.label_smoothing / (C - 1)
            const = on * math.log(on) + (C - 1) * off * math.log(off)
            ce = const - on * logp_y - off * (logp.sum(-1) - logp_y)
        else:
            ce = -logp_y
        if mask is not None:
            m = mask.float()
            loss_ss = (ce * m).sum() / m.sum().clamp_min(1.0)
        else:
            loss_ss = ce.mean()

        ss_tv = zero
        if self.ss_tv_lambda > 0.0 and logits.size(1) >= 2:
            p = F.softmax(logits, dim=-1)
            tv = (p[:, 1:] - p[:, :-1]).abs().sum(-1)
            if mask is not None:
                tm = (mask[:, 1:] & mask[:, :-1]).float()
                ss_tv = (tv * tm).sum() / tm.sum().clamp_min(1.0)
            else:
                ss_tv = tv.mean()

        if self.use_vq and self.quantizer is not None:       # models/vq_vae.py:1292-1294
            vq_loss = self.quantizer.beta * self.quantizer.commitment_loss(zq_raw, ze_raw)
        else:
            vq_loss = zero

        usage_reg = zero                                     # models/vq_vae.py:1298-1309
        if self.usage_entropy_lambda > 0.0 and ze_raw.numel() > 0 and self.quantizer is not None:
            p_code = self.quantizer.usage_code_probs(ze_raw)
            entropy = -(p_code * p_code.clamp_min(1e-12).log()).sum()
            usage_reg = -self.usage_entropy_lambda * entropy

        total = rmsd_weight * loss_xyz + ss_weight * loss_ss + vq_loss + self.ss_tv_lambda * ss_tv + usage_reg
        with torch.no_grad():
            hit = logits.argmax(-1) == labels
            acc = (hit & mask).sum().float() / mask.sum().float().clamp_min(1.0) if mask is not None else hit.float().mean()
        det = lambda t: t.detach()
        return {"loss": total, "Reconstruction_Loss_XYZ": det(loss_xyz), "XYZ_MSE_Raw": det(loss_xyz),
                "XYZ_MSE_Aligned": det(loss_xyz), "Reconstruction_Loss_SS": det(loss_ss), "SS_Accuracy": det(acc),
                "VQ_Loss": det(vq_loss), "Geom_BondLength_Loss": zero, "Geom_BondAngle_Loss": zero,
                "Geom_Direction_Loss": zero, "Geom_Dihedral_Loss": zero, "Geom_Loss": zero, "SS_TV": det(ss_tv),
                "Usage_Reg": det(usage_reg), "XYZ_TV2": zero, "VQ_Perplexity": det(ppl), "VQ_DeadRatio": det(dead),
                "RMSD_Raw": rmsd, "RMSD_Aligned": rmsd}

    @torch.no_grad()
    def generate(self, x: Tensor, mask: Optional[Tensor] = None, **kwargs):
        return self.forward(x, mask=mask)[0]

    @torch.no_grad()
    def sample(self, num_samples: int, device, out_len: Optional[int] = None):
        """Random codes -> latent -> decode (models/vq_vae.py:1394-1422).  The per-level gather + level sum
        runs in the fused indices_to_latent kernel on token-major ids."""
        if not self.use_vq or self.quantizer is None:
            raise RuntimeError("Quantizer is not initialized for sampling.")
        q, M = self.quantizer, int(self.latent_n_tokens)
        L_out = out_len if out_len is not None else self.max_seq_len
        Q = q.num_quantizers if self.residual_vq else 1
        K_pick = q.K_per if self.residual_vq else q.K
        idx = torch.randint(0, K_pick, (num_samples, M, Q), device=device)
        if Q > 1:
            idx = idx + torch.arange(Q, device=device) * q.K_per                 # global ids, token-major [B, M, Q]
        z_q = ops.indices_to_latent(idx.reshape(-1), q.embedding, Q).view(num_samples, M, -1)
        return self.decode(z_q, mask=torch.ones(num_samples, L_out, dtype=torch.bool, device=device))
