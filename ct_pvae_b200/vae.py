"""Torch restatement of the CALLERS of the Radon path, just enough
to run CT_PVAE's training step end to end on the new projector and report it/s.

TensorFlow is not installable in this image, so the reference driver itself cannot
execute here; this module restates, without the projector arithmetic (that is
libctradon's job):

  ctvae/models.py:23-342          create_encode_net / create_decode_net / conv_block (maxout) / periodic_padding
  ctvae/helper_functions.py:198   positive_range
  ctvae/helper_functions.py:204   find_loss_vae_unsup   (ELBO: KL - E[log p(M|R) + log p(R|z)])
  ctvae/main_ct_vae.py:463-486    train_step            (loss/1e5, NaN scrub, per-tensor clip_by_norm, Adam)
  ctvae/create_masks.py:16-107    create_all_masks      (sparse-angle masks, Poisson-noised sinograms)
  ctvae/helper_functions.py:477   iradon_all            (initial reconstructions; FBP via ct_pvae_b200.iradon
                                                        instead of tomopy gridrec, which is not installable)
  ctvae/helper_functions.py:33    create_sinogram       (ct_pvae_b200.project_tf_fast instead of tomopy.project)

The dense conv nets run on cuDNN through torch (they are out of the hot path's scope);
the measurement term goes through the fused log-likelihood kernel.  Tensors are NCHW
here (the reference is NHWC); weights are random-initialised (Glorot), so this is a
throughput and plumbing vehicle, not a weight-compatible port.
"""
from __future__ import annotations

import math
from typing import List, Sequence

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .fbp_tensorflow import get_fourier_filter, iradon
from .forward_functions import project_tf_fast
from .likelihood import log_prob_M_given_R_sum

EPS32 = float(np.finfo(np.float32).eps)


def positive_range(x: torch.Tensor, offset: float = EPS32) -> torch.Tensor:
    """helper_functions.py:198-201: exp(x-1)+eps below 1, identity above."""
    xm = x - 1
    neg = (xm < 0).to(x.dtype)
    return (torch.exp(torch.clamp(xm, -1e10, 10)) + offset) * neg + (xm + 1) * (1 - neg)


def periodic_padding(x: torch.Tensor, pad_x: Sequence[int], pad_y: Sequence[int]) -> torch.Tensor:
    """models.py:219-263 (wrap-around padding of rows then columns), NCHW."""
    return F.pad(x, (pad_y[0], pad_y[1], pad_x[0], pad_x[1]), mode="circular") if any(pad_x) or any(pad_y) else x


class ConvBlock(nn.Module):
    """models.py:267-342: (dropout) -> two convs -> maxout -> (norm).  The two convs are
    one conv with twice the channels followed by a max over the halves."""

    def __init__(self, cin: int, cout: int, kernel: int, stride: int = 2, transpose: bool = False, dropout: float = 0.0):
        super().__init__()
        self.k, self.s, self.transpose, self.cout = kernel, stride, transpose, cout
        self.drop = nn.Dropout(dropout) if dropout > 0 else None
        if transpose:
            if kernel < stride:
                raise ValueError("transposed conv_block needs kernel_size >= stride")
            # Keras Conv2DTranspose(padding="same") (models.py:287-301): output = input * stride.  The full transposed
            # convolution is (input - 1) * stride + kernel long; "same" drops kernel - stride samples, the smaller half
            # in front (the toy run's kernel 2 / stride 1 drops one sample at the end)
            self.conv = nn.ConvTranspose2d(cin, 2 * cout, kernel, stride=stride, padding=0)
            self.crop = (kernel - stride) // 2
        else:
            self.conv = nn.Conv2d(cin, 2 * cout, kernel, stride=stride)
        nn.init.xavier_uniform_(self.conv.weight)
        nn.init.zeros_(self.conv.bias)

    def forward(self, x):
        if self.drop is not None:
            x = self.drop(x)
        if not self.transpose:
            hx, hy = x.shape[-2], x.shape[-1]
            px = self.k - (hx % self.s if hx % self.s else self.s)
            py = self.k - (hy % self.s if hy % self.s else self.s)
            x = periodic_padding(x, (px // 2 + px % 2, px // 2), (py // 2 + py % 2, py // 2))
        y = self.conv(x)
        if self.transpose:
            hx, hy = x.shape[-2] * self.s, x.shape[-1] * self.s
            y = y[..., self.crop:self.crop + hx, self.crop:self.crop + hy]
        return torch.maximum(y[:, :self.cout], y[:, self.cout:])


class EncodeNet(nn.Module):
    """models.py:23-108: input [B, num_filters+1, X, Y] -> list of skips (num_blocks+1)."""

    def __init__(self, in_ch: int, num_feature_maps_vec, num_blocks=3, kernel_size=4, stride_encode=2,
                 intermediate_layers=2, intermediate_kernel=4, feature_maps_multiplier=2, dropout=0.0):
        super().__init__()
        self.fmm = feature_maps_multiplier
        ch = in_ch * feature_maps_multiplier
        self.skip_channels = [ch]
        self.blocks = nn.ModuleList()
        for i in range(num_blocks):
            layers = [ConvBlock(ch, ch, intermediate_kernel, 1, dropout=dropout) for _ in range(intermediate_layers)]
            out = int(feature_maps_multiplier * num_feature_maps_vec[i])
            layers.append(ConvBlock(ch, out, kernel_size, stride_encode, dropout=dropout))
            self.blocks.append(nn.Sequential(*layers))
            ch = out
            self.skip_channels.append(ch)

    def forward(self, x):
        x = x.repeat_interleave(self.fmm, dim=1)
        skips = [x]
        for blk in self.blocks:
            x = blk(x)
            skips.append(x)
        return skips


class DecodeNet(nn.Module):
    """models.py:112-215: sampled skips (channels / fmm) -> (alpha, beta) maps [B,1,X,Y]."""

    def __init__(self, skip_channels, final_output_channels=1, kernel_size=4, stride_encode=2, intermediate_layers=2,
                 intermediate_kernel=4, feature_maps_multiplier=2, dropout=0.0):
        super().__init__()
        zin = [c // feature_maps_multiplier for c in skip_channels]
        self.ups = nn.ModuleList()
        ch = zin[-1]
        for lvl in range(len(skip_channels) - 2, -1, -1):
            out = skip_channels[lvl]
            layers = [ConvBlock(ch, out, kernel_size, stride_encode, transpose=True, dropout=dropout)]
            layers += [ConvBlock(out, out, intermediate_kernel, 1, dropout=dropout) for _ in range(intermediate_layers)]
            self.ups.append(nn.Sequential(*layers))
            ch = out + zin[lvl]
        self.final = ConvBlock(ch, 2 * final_output_channels, kernel_size, 1, dropout=dropout)
        self.nout = final_output_channels

    def forward(self, skips: List[torch.Tensor]):
        out = skips[-1]
        for up, skip in zip(self.ups, reversed(skips[:-1])):
            out = up(out)
            rx, ry = out.shape[-2] - skip.shape[-2], out.shape[-1] - skip.shape[-1]
            ox, oy = rx // 2 + rx % 2, ry // 2 + ry % 2
            out = out[..., ox:ox + skip.shape[-2], oy:oy + skip.shape[-1]]
            out = torch.cat([out, skip], dim=1)
        out = self.final(out)
        return out[:, :self.nout], out[:, self.nout:]


# ---------------------------------------------------------------------------------- distributions
_SQRT2 = math.sqrt(2.0)


def _ndtr(x):
    return 0.5 * (1 + torch.erf(x / _SQRT2))


class TruncatedNormal:
    """tfd.TruncatedNormal(loc, scale, low, high) with a reparameterised sample."""

    def __init__(self, loc, scale, low=0.0, high=1e10):
        self.loc, self.scale, self.low, self.high = loc, scale, low, high
        self.a, self.b = (low - loc) / scale, (high - loc) / scale
        self.cdf_a, self.cdf_b = _ndtr(self.a), _ndtr(self.b)
        self.z = torch.clamp(self.cdf_b - self.cdf_a, min=1e-30)

    def sample(self):
        u = torch.rand_like(self.loc)
        p = torch.clamp(self.cdf_a + u * self.z, 1e-7, 1 - 1e-7)
        x = self.loc + self.scale * torch.special.ndtri(p)
        return torch.clamp(x, min=self.low)

    def log_prob(self, x):
        zz = (x - self.loc) / self.scale
        return -0.5 * zz * zz - torch.log(self.scale) - 0.5 * math.log(2 * math.pi) - torch.log(self.z)

    def mean(self):
        pa = torch.exp(-0.5 * self.a ** 2) / math.sqrt(2 * math.pi)
        pb = torch.exp(-0.5 * self.b ** 2) / math.sqrt(2 * math.pi)
        return self.loc + self.scale * (pa - pb) / self.z


# ---------------------------------------------------------------------------------- loss and step
def find_loss_vae_unsup(proj_sample, mask, input_encode, model_encode, model_decode, poisson_noise_multiplier, sqrt_reg,
                        kl_anneal=1.0, kl_multiplier=1.0, num_samples=2, theta=None, angles_i=None, pad=True,
                        use_normal=True, training=True, interpolation="nearest", adjoint="exact"):
    """helper_functions.py:204-332.  input_encode [B,C,X,Y] (NCHW), mask [B,A], proj_sample [B,A,P].
    The distributions are built with validate_args=False: the argument checks are host-synchronising reductions
    (``(scale > 0).all()``), which stall the launch queue every call and cannot run under CUDA-graph capture."""
    skips_val = model_encode(input_encode / 300)
    q = []
    for sv in skips_val:
        loc, log_scale = sv.chunk(2, dim=1)
        scale = positive_range(log_scale)
        q.append(torch.distributions.Normal(loc, scale + sqrt_reg, validate_args=False) if use_normal
                 else torch.distributions.Beta(positive_range(loc), scale, validate_args=False))
    log_prob_M = []
    out_dists = []
    for _ in range(num_samples):
        q_sample = [d.rsample() for d in q]
        alpha, beta = model_decode(q_sample)
        if use_normal:
            out = TruncatedNormal(positive_range(alpha), positive_range(beta), 0.0, 1e10)
            x = out.sample()
            lp_R = out.log_prob(x)
        else:
            out = torch.distributions.Beta(positive_range(alpha), positive_range(beta), validate_args=False)
            x = out.rsample()
            lp_R = out.log_prob(torch.clamp(x, sqrt_reg, 1 - sqrt_reg))
        out_dists.append(out)
        # [B,1,X,Y] -> the projector's [B,X,Y,1]; fused projector + log p(M|R) + reduction
        lp_M = log_prob_M_given_R_sum(x.permute(0, 2, 3, 1), mask, proj_sample, poisson_noise_multiplier, sqrt_reg,
                                      theta=theta, angles_i=angles_i, pad=pad, interpolation=interpolation, adjoint=adjoint)
        log_prob_M.append(lp_M + lp_R.sum())
    if use_normal:
        # (ones_like, not the Python scalar 1.0: torch would build the scalar on the host and copy it to the device --
        # a host->device copy per call, and illegal under CUDA-graph capture)
        prior = [torch.distributions.Normal(torch.zeros_like(d.loc), torch.ones_like(d.loc), validate_args=False) for d in q]
    else:
        prior = [torch.distributions.Beta(torch.full_like(d.concentration1, 0.5), torch.full_like(d.concentration0, 0.5), validate_args=False) for d in q]
    kl = sum(torch.distributions.kl_divergence(q[i], prior[i]).sum(dim=(1, 2, 3)) for i in range(1, len(q)))
    loglik = torch.stack(log_prob_M).mean(dim=0)
    return kl_anneal * kl_multiplier * kl - loglik, out_dists, kl, loglik


class CTVAE(nn.Module):
    """Networks + optimiser of main_ct_vae.py:258-373 with the reference's CLI defaults."""

    def __init__(self, x_size, y_size, num_filters=1, num_feature_maps=20, num_feature_maps_multiplier=1.1, num_blocks=3,
                 kernel_size=4, stride_encode=2, intermediate_layers=2, intermediate_kernel=4, dropout_prob=0.0,
                 learning_rate=1e-4, adam_epsilon=1e-7):
        super().__init__()
        fmv = [int(num_feature_maps * num_feature_maps_multiplier ** i) for i in range(num_blocks)]
        self.encode = EncodeNet(num_filters + 1, fmv, num_blocks, kernel_size, stride_encode, intermediate_layers,
                                intermediate_kernel, 2, dropout_prob)
        self.decode = DecodeNet(self.encode.skip_channels, 1, kernel_size, stride_encode, intermediate_layers,
                                intermediate_kernel, 2, dropout_prob)
        self.x_size, self.y_size = x_size, y_size
        self.lr, self.adam_eps = learning_rate, adam_epsilon
        self.optimizer = None

    def train_step(self, proj_sample, mask, input_encode, pnm, theta, angles_i=None, pad=True, num_samples=2, norm=100.0,
                   kl_anneal=1.0, kl_multiplier=1.0, use_normal=True, sqrt_reg=EPS32, training=True, interpolation="nearest"):
        """main_ct_vae.py:463-486."""
        if self.optimizer is None:
            # capturable: the step counter lives on the device, so the optimiser step can be part of a CUDA graph
            cuda = next(self.parameters()).is_cuda
            self.optimizer = torch.optim.Adam(self.parameters(), lr=self.lr, eps=self.adam_eps, capturable=cuda)
        with torch.set_grad_enabled(training):
            loss, out_dists, kl, loglik = find_loss_vae_unsup(
                proj_sample, mask, input_encode, self.encode, self.decode, pnm, sqrt_reg, kl_anneal, kl_multiplier,
                num_samples, theta, angles_i, pad, use_normal, training, interpolation)
            loss = loss.mean() / 1e5
        if training:
            self.optimizer.zero_grad(set_to_none=True)
            loss.backward()
            self._average_gradients()
            for p in self.parameters():
                if p.grad is not None:
                    torch.nan_to_num_(p.grad, nan=0.0)
                    # tf.clip_by_norm per tensor: t * clip / max(||t||, clip), without reading the norm on the host
                    p.grad.mul_(norm / torch.clamp(p.grad.norm(), min=norm))
            self.optimizer.step()
        return loss.detach(), out_dists, kl.detach(), loglik.detach()


    def _average_gradients(self):
        """Batch-sharded data parallelism (BASELINE configs[2]): every rank holds its own slice of
        the batch; one flat all-reduce averages the network gradients.  The projector itself
        needs no collective in this mode."""
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        grads = [p.grad for p in self.parameters() if p.grad is not None]
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat)
        flat /= dist.get_world_size()
        off = 0
        for g in grads:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n


class GraphedTrainStep:
    """main_ct_vae.py:375-422's iteration as ONE CUDA graph: batch gather, encoder, ``num_samples`` decoder passes, the
    fused projector + log-likelihood (forward and adjoint kernels), backward, gradient clipping and the Adam step are
    captured once and replayed.  The eager step issues ~2200 small launches per iteration and is bound by the host's
    launch rate (profiles/r2_prof_train_before_graphs.txt: 21 ms per iteration for 8.7 ms of GPU work); a replay
    costs one launch.  The per-iteration inputs -- the example indices and the angle minibatch (main_ct_vae.py:388-389)
    -- are written into static device tensors before each replay; the angle subset reaches the kernels as an index
    list into ONE plan (ctr_radon_loglik_sel), so nothing on the iteration path allocates or creates a plan."""

    def __init__(self, model: "CTVAE", proj_samples, masks, input_encode, pnm, theta, batch: int, angles_per_iter: int,
                 num_samples: int = 2, pad: bool = True, interpolation: str = "nearest", warmup: int = 3, **step_kwargs):
        dev = proj_samples.device
        if dev.type != "cuda":
            raise RuntimeError("GraphedTrainStep needs CUDA tensors")
        self.model, self.data = model, (proj_samples, masks, input_encode)
        self.idx = torch.zeros((batch,), dtype=torch.long, device=dev)
        self.sel = torch.arange(angles_per_iter, dtype=torch.int32, device=dev)
        self.args = dict(pnm=pnm, theta=theta, pad=pad, num_samples=num_samples, interpolation=interpolation, **step_kwargs)
        self.graph = None
        self.loss = None
        # warm-up on a side stream (allocator, cuDNN algorithm choice, lazy optimiser state), then capture
        s = torch.cuda.Stream(dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(max(1, warmup)):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        model.optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(g):
            self.loss = self._body()
        self.graph = g

    def _body(self):
        meas, masks, enc = self.data
        loss, _, _, _ = self.model.train_step(meas[self.idx], masks[self.idx], enc[self.idx], self.args["pnm"], self.args["theta"],
                                              angles_i=self.sel, **{k: v for k, v in self.args.items() if k not in ("pnm", "theta")})
        return loss

    def __call__(self, example_idx, angles_i):
        """One training iteration on the examples ``example_idx`` [batch] at the angles ``angles_i`` [angles_per_iter];
        returns the (device) loss of this iteration."""
        self.idx.copy_(torch.as_tensor(example_idx), non_blocking=True)
        self.sel.copy_(torch.as_tensor(angles_i).to(torch.int32), non_blocking=True)
        self.graph.replay()
        return self.loss


# ---------------------------------------------------------------------------------- data preparation
def create_sinogram(imgs, theta, pad=True, interpolation="bilinear"):
    """helper_functions.py:33-38 with the new projector instead of tomopy.project:
    imgs [N,X,Y] -> sinograms [N,A,P].  (tomopy's projector is a different discretisation.)"""
    t = torch.as_tensor(imgs)
    return project_tf_fast(t.unsqueeze(-1), theta, pad=pad, dim=2, integrate_vae=True, interpolation=interpolation)[..., 0]


def create_all_masks(x_train_sinograms, num_angles, poisson_noise_multiplier=1e3, num_sparse_angles=10, random=False,
                     toy_masks=False, generator=None):
    """create_masks.py:16-107: masks [N,A] with value 1/nsa at the kept angles, and the
    Poisson-noised masked sinograms [N,A,P]."""
    s = torch.clamp(torch.as_tensor(x_train_sinograms, dtype=torch.float32), min=0)
    n = s.shape[0]
    if toy_masks:
        masks = torch.tensor([[1, 0], [0, 1], [1, 0], [0, 1]], dtype=torch.float32).repeat(n // 4, 1)
    else:
        masks = torch.zeros((n, num_angles))
        for ind in range(n):
            if random:
                sel = torch.randperm(num_angles, generator=generator)[:num_sparse_angles]
            else:
                spacing = math.ceil(num_angles / num_sparse_angles)
                sel = (torch.arange(0, spacing * num_sparse_angles, spacing) % num_angles).long()
            masks[ind].index_add_(0, sel, torch.ones(len(sel)))
        masks = masks / num_sparse_angles
    masks = masks.to(s.device)
    proj_masked = s * masks[:, :, None]
    proj_samples = torch.poisson(proj_masked * poisson_noise_multiplier, generator=generator) / poisson_noise_multiplier
    return masks, proj_samples


def iradon_all(all_proj_samples, all_masks, theta, x_size, y_size, sqrt_reg=EPS32, filter_name="ramp"):
    """helper_functions.py:477-529 with FBP (ct_pvae_b200.iradon) standing in for tomopy:
    channel 0 = reconstruction of the dose-normalised masked sinogram, channel 1 = unfiltered
    back-projection of the mask.  Returns [N,2,X,Y] (NCHW)."""
    P = all_proj_samples.shape[-1]
    m = all_masks[:, :, None].expand(-1, -1, P)
    ps = torch.where(m > sqrt_reg, all_proj_samples / torch.clamp(m, min=sqrt_reg), all_proj_samples)
    rec = iradon(ps, theta, x_size, y_size, get_fourier_filter(P, filter_name))
    rec_mask = iradon(m.contiguous(), theta, x_size, y_size, get_fourier_filter(P, None))
    return torch.stack([rec, rec_mask], dim=1).float()
