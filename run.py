"""Entry points of the reference's run.py for the hot path: --mode train / eval_fid / save_latent /
train_latent_ddim / eval, same option names and defaults (reference run.py:25-97), same folder and file naming
(models/<exp>/model-N.pth, imgs/<exp>/eval-fid-*/sample-%06d.png, <model>_<exp>_latent.npz), on the sm_100a kernels.

What differs, on purpose:
  * data: the reference downloads torchvision datasets (data.py); there is no network here and data loading is out
    of scope, so images come from --data_npz (array `images`: N x H x W x C uint8 or N x C x H x W
    float in [-1, 1], optional `labels`) or are synthetic U[-1, 1] (--synthetic_size images);
  * shapes: the kernels cover the 64x64 / 64-channel configuration (celeba, ffhq, chairs-sized data with
    --unets_channels 64); other dataset names raise;
  * torchrun: under WORLD_SIZE > 1 training is data parallel (--batch_size per GPU, like BASELINE configs[2]);
    sampling / encoding shard every round of --batch_size images over the ranks, all ranks drawing the round's full
    batch from the same seed, so the written images do not depend on the number of GPUs (checked by
    tools/check_run_py_multigpu.sh: 2 GPUs vs 1 GPU differ by at most one uint8 level in 0.01 % of the pixels).
The VAE baseline and the analysis modes (disentangle, interpolate, latent_quality, plot_latent) are not built.
"""
from __future__ import annotations

import argparse
import os
import random

import numpy as np
import torch

from infodiffusion_b200 import io as idf_io
from infodiffusion_b200.distributed import gather_batch, local_slice
from infodiffusion_b200.layout import shard_range
from infodiffusion_b200.models import Diff, InfoDiff
from infodiffusion_b200.optim import ClipAdamW
from infodiffusion_b200.sampling import DiffusionProcess, LatentDiffusionProcess, TwoPhaseDiffusionProcess
from infodiffusion_b200.train import GradSync, set_grad_sync


def parse_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument('--r_seed', type=int, default=0)
    p.add_argument('--img_id', type=int, default=0)
    p.add_argument('--model', required=True, choices=['diff', 'vae', 'vanilla'])
    p.add_argument('--mode', required=True, choices=['train', 'eval', 'eval_fid', 'save_latent', 'disentangle', 'interpolate',
                                                     'save_original_img', 'latent_quality', 'train_latent_ddim', 'plot_latent'])
    p.add_argument('--prior', required=True, choices=['regular', '10mix', 'roll'])
    p.add_argument('--kld_weight', type=float, default=0)
    p.add_argument('--mmd_weight', type=float, default=0.1)
    p.add_argument('--use_C', action='store_true', default=False)
    p.add_argument('--C_max', type=float, default=25)
    p.add_argument('--dataset', required=True, choices=['fmnist', 'mnist', 'celeba', 'cifar10', 'dsprites', 'chairs', 'ffhq',
                                                        'synthetic'])
    p.add_argument('--img_folder', default='./imgs')
    p.add_argument('--log_folder', default='./logs')
    p.add_argument('-e', '--epochs', type=int, default=20)
    p.add_argument('--save_epochs', type=int, default=5)
    p.add_argument('--batch_size', type=int, default=64)
    p.add_argument('--learning_rate', type=float, default=0.0001)
    p.add_argument('--optimizer', default='adam', choices=['adam'])
    p.add_argument('--model_folder', default='./models')
    p.add_argument('--deterministic', action='store_true', default=False)
    p.add_argument('--input_channels', type=int, default=1)
    p.add_argument('--unets_channels', type=int, default=64)
    p.add_argument('--encoder_channels', type=int, default=64)
    p.add_argument('--input_size', type=int, default=32)
    p.add_argument('--a_dim', type=int, default=32, required=True)
    p.add_argument('--beta1', type=float, default=1e-5)
    p.add_argument('--betaT', type=float, default=1e-2)
    p.add_argument('--diffusion_steps', type=int, default=1000)
    p.add_argument('--split_step', type=int, default=500)
    p.add_argument('--sampling_number', type=int, default=16)
    p.add_argument('--data_dir', type=str, default='./data')
    p.add_argument('--tb_logger', action='store_true')
    p.add_argument('--is_latent', action='store_true')
    p.add_argument('--is_bottleneck', action='store_true')
    # additions (see module docstring)
    p.add_argument('--data_npz', type=str, default=None, help='images (and labels) to train on / encode')
    p.add_argument('--synthetic_size', type=int, default=256, help='number of synthetic images when no --data_npz is given')
    return p.parse_args(argv)


def generate_exp_string(args) -> str:
    """reference utils.py:49-61"""
    root = f'{args.dataset}_{args.a_dim}d'
    if args.kld_weight != 0:
        root += f'_{args.kld_weight}kld'
        if args.use_C:
            root += f'_{args.C_max}C'
    if args.mmd_weight != 0:
        root += f'_{args.mmd_weight}mmd'
    if args.prior != 'regular':
        root += f'_{args.prior}'
    if args.is_bottleneck:
        root += '_bottleneck'
    return root


def seed_everything(r_seed: int) -> None:
    """reference utils.py:64-71"""
    random.seed(r_seed)
    np.random.seed(r_seed)
    torch.manual_seed(r_seed)
    torch.cuda.manual_seed_all(r_seed)


def get_dataset_config(args):
    """reference data.py:63-108, restricted to what the kernels cover."""
    if args.dataset in ('celeba', 'ffhq', 'synthetic'):
        args.input_channels, args.unets_channels, args.encoder_channels, args.input_size = 3, 64, 64, 64
    elif args.dataset == 'cifar10':
        args.input_channels, args.unets_channels, args.encoder_channels, args.input_size = 3, 64, 64, 32
    else:
        raise NotImplementedError(f"--dataset {args.dataset}: the sm_100a plans cover 64x64 / 32x32 inputs with 64-channel "
                                  "UNets (celeba / ffhq / cifar10 shapes); see DESIGN.md section 7")
    return (args.input_channels, args.input_size, args.input_size)


def _world():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not torch.distributed.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        torch.distributed.init_process_group("nccl")
    return rank, world


def load_images(args, shape):
    """[N, C, H, W] fp32 in [-1, 1] on the CPU (+ labels or None)."""
    if args.data_npz:
        z = np.load(args.data_npz)
        x = z["images"]
        x = torch.from_numpy(x)
        if x.dtype == torch.uint8:
            x = x.permute(0, 3, 1, 2).float() / 127.5 - 1.0
        labels = z["labels"] if "labels" in z.files else None
    else:
        g = torch.Generator().manual_seed(args.r_seed)
        x = torch.rand(args.synthetic_size, *shape, generator=g) * 2 - 1
        labels = None
    assert tuple(x.shape[1:]) == tuple(shape), f"data shape {tuple(x.shape[1:])} != {shape}"
    return x.float(), labels


class GradualWarmupScheduler(torch.optim.lr_scheduler._LRScheduler):
    """reference utils.py:133-160 (linear warm-up to multiplier x base lr over warm_epoch epochs, then the wrapped
    scheduler)."""

    def __init__(self, optimizer, multiplier, warm_epoch, after_scheduler=None):
        self.multiplier, self.total_epoch, self.after_scheduler, self.finished = multiplier, warm_epoch, after_scheduler, False
        super().__init__(optimizer)

    def get_lr(self):
        if self.last_epoch > self.total_epoch:
            if self.after_scheduler:
                if not self.finished:
                    self.after_scheduler.base_lrs = [b * self.multiplier for b in self.base_lrs]
                    self.finished = True
                return self.after_scheduler.get_lr()
            return [b * self.multiplier for b in self.base_lrs]
        return [b * ((self.multiplier - 1.) * self.last_epoch / self.total_epoch + 1.) for b in self.base_lrs]

    def step(self, epoch=None):
        if self.finished and self.after_scheduler:
            self.after_scheduler.step(None if epoch is None else epoch - self.total_epoch)
        else:
            super().step(epoch)


def _model_root(args, latent: bool = False) -> str:
    root = args.model_folder
    if args.model == 'vanilla':
        root = os.path.join(root, 'diff')
    root = os.path.join(root, generate_exp_string(args))
    return root + ('_latent' if latent else '')


def _fit(args, model, batches, device, rank, world, latent=False):
    """The reference's epoch loop (run.py:187-211 / 503-526): loss_fn -> backward -> clip(1.0) -> AdamW, cosine schedule
    with one warm-up epoch, checkpoint every save_epochs."""
    opt = ClipAdamW(model.parameters(), lr=args.learning_rate, weight_decay=1e-5, max_norm=1.0)
    cosine = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer=opt, T_max=args.epochs, eta_min=0, last_epoch=-1)
    warm = GradualWarmupScheduler(optimizer=opt, multiplier=2., warm_epoch=1, after_scheduler=cosine)
    sync = GradSync(world) if world > 1 else None
    set_grad_sync(sync)
    params = [p for p in model.parameters() if p.requires_grad]
    model.train()
    for epoch in range(args.epochs):
        total, n = torch.zeros((), device=device), 0
        for data in batches(epoch):
            loss = model.loss_fn(args=args, x=data.to(device, non_blocking=True), curr_epoch=epoch)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            if sync is not None:
                sync.finish(params)
            opt.step()
            total += loss.detach()                      # no host sync inside the epoch
            n += 1
        if rank == 0:
            print(f"Epoch [{epoch}/{args.epochs}] Loss {float(total) / max(n, 1):.4f}  lr {opt.param_groups[0]['lr']:.3e}")
        warm.step()
        if (epoch + 1) % args.save_epochs == 0 and rank == 0:
            root = _model_root(args, latent)
            os.makedirs(root, exist_ok=True)
            torch.save(model.state_dict(), os.path.join(root, f'model-{epoch + 1}.pth'))
    set_grad_sync(None)
    model.eval()


def _batches(x, batch_size, rank, world, seed):
    def it(epoch):
        g = torch.Generator().manual_seed(seed * 1000 + epoch)
        perm = torch.randperm(x.shape[0], generator=g)
        usable = (x.shape[0] // (batch_size * world)) * batch_size * world      # fixed batch shape: plans are per batch size
        for i in range(0, usable, batch_size * world):
            idx = perm[i + rank * batch_size: i + (rank + 1) * batch_size]
            yield x[idx].pin_memory()
    return it


def train(args):
    rank, world = _world()
    seed_everything(args.r_seed)
    device = torch.device("cuda", torch.cuda.current_device())
    shape = get_dataset_config(args)
    x, _ = load_images(args, shape)
    model = InfoDiff(args, device, shape) if args.model == 'diff' else Diff(args, device, shape)
    _fit(args, model, _batches(x, args.batch_size, rank, world, args.r_seed), device, rank, world)


def train_latent_ddim(args):
    rank, world = _world()
    seed_everything(args.r_seed)
    device = torch.device("cuda", torch.cuda.current_device())
    z = torch.from_numpy(np.load("{}_{}_latent.npz".format(args.model, generate_exp_string(args).replace(".", "_")))["all_a"]).float()
    model = Diff(args, device, (1, args.a_dim, args.a_dim))
    _fit(args, model, _batches(z, args.batch_size, rank, world, args.r_seed), device, rank, world, latent=True)


def _load(args, device, shape):
    model = InfoDiff(args, device, shape) if args.model == 'diff' else Diff(args, device, shape)
    path = os.path.join(_model_root(args), f'model-{args.epochs}.pth')
    if os.path.exists(path):
        model.load_state_dict(torch.load(path, map_location=device), strict=False)
    elif int(os.environ.get("RANK", "0")) == 0:
        print(f"[run.py] no checkpoint at {path}: using the seeded random initialisation")
    return model.eval()


def evaluate(args):
    rank, world = _world()
    seed_everything(args.r_seed)
    device = torch.device("cuda", torch.cuda.current_device())
    shape = get_dataset_config(args)
    model = _load(args, device, shape)
    exp = generate_exp_string(args)
    if args.mode in ('eval', 'eval_fid'):
        process = DiffusionProcess(args, model, device, shape)
        process_latent = None
        if args.mode == 'eval_fid' and args.model == 'diff':
            if args.is_latent:                                   # run.py:234-243, 278
                model2 = Diff(args, device, (1, args.a_dim, args.a_dim))
                p2 = f'./models/{exp}_latent/model-{args.epochs}.pth'
                if not os.path.exists(p2):
                    raise FileNotFoundError(f"The file path {p2} does not exist, please train the latent diffusion model first.")
                model2.load_state_dict(torch.load(p2, map_location=device), strict=True)
                process_latent = LatentDiffusionProcess(args, model2.eval(), device)
            else:                                                # run.py:244-252, 280: two-phase with the vanilla model
                p2 = f'./models/diff/{args.dataset}_{args.a_dim}d/model-{args.epochs}.pth'
                if os.path.exists(p2):
                    model2 = Diff(args, device, shape)
                    model2.load_state_dict(torch.load(p2, map_location=device), strict=True)
                    process = TwoPhaseDiffusionProcess(args, model, model2.eval(), device, shape)
        # run.py:265-274 (eval_fid: imgs/<exp>/eval-fid-latent|eval-fid-fast) and save_images (eval: imgs[/diff]/<exp>/eval)
        if args.mode == 'eval_fid':
            root = os.path.join(args.img_folder, exp, 'eval-fid-latent' if args.is_latent else 'eval-fid-fast')
        else:
            root = os.path.join(args.img_folder, 'diff' if args.model == 'vanilla' else '', exp, 'eval')
        total = args.sampling_number
        for first in range(0, total, args.batch_size):               # --batch_size is the GLOBAL batch of a round
            n_round = min(args.batch_size, total - first)
            lo, hi = shard_range(n_round, rank, world)
            # every rank draws the round's FULL batch from the same seed (x_T, then a, on the CPU generator like
            # sampling.py:92-95; per-step noise on the CUDA generator) and keeps its slice: the images do not depend
            # on the number of GPUs, and with one GPU the draws are the reference's
            if process_latent is not None:
                a_full = process_latent.sampling(sampling_number=n_round)
            else:
                a_full = None
            xT_full = torch.randn([n_round, *shape]).to(device=device)
            if a_full is None:
                a_full = torch.randn([n_round, args.a_dim]).to(device=device)
            if hi == lo:
                continue
            if world > 1:
                noise = lambda idx, out: out.copy_(torch.randn(n_round, *shape, device=device)[lo:hi])
                for pr in (process, getattr(process, "p1", None), getattr(process, "p2", None)):
                    if pr is not None:
                        pr.noise_fn = noise
            batch = process.sampling(sampling_number=hi - lo, xT=xT_full[lo:hi].contiguous(), a=a_full[lo:hi].contiguous())
            idf_io.save_eval_images(batch, root, first_index=first + lo, limit=total)     # every rank writes its own range
        if rank == 0:
            print("DONE", root)
    elif args.mode == 'save_latent':
        x, labels = load_images(args, shape)
        outs = []
        for i in range(0, x.shape[0], args.batch_size * world):
            chunk = x[i:i + args.batch_size * world]
            part = local_slice(chunk, chunk.shape[0]).to(device)
            a, _, mu, _ = model.encoder(part) if part.shape[0] else (torch.zeros(0, args.a_dim, device=device),) * 4
            z = mu if args.kld_weight != 0 else a                # run.py:428-437
            outs.append(gather_batch(z.contiguous(), chunk.shape[0]).cpu())
        if rank == 0:
            attr = labels if labels is not None else np.array(['No Attributes'] * x.shape[0])
            idf_io.save_latents_npz("{}_{}_latent".format(args.model, exp.replace(".", "_")), outs, [attr])
    else:
        raise NotImplementedError(f"--mode {args.mode} is an analysis mode outside the hot path (SURVEY section 8f rank 4)")


if __name__ == '__main__':
    args = parse_args()
    if args.model == 'vae':
        raise NotImplementedError("the VAE baseline is out of scope (SURVEY section 2)")
    if args.mode == 'train':
        train(args)
    elif args.mode == 'train_latent_ddim':
        train_latent_ddim(args)
    else:
        evaluate(args)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()
