"""Drop-in for the reference's train.py: same 27 flags, same Train_GAN surface, same files written
(models/<folder_save>/{final_model.pth, params.txt, *loss.npy}, checkpoints/<folder_save>/model_<e>.pth).

The loop body (reference train.py:99-168) is one call to step.TrainStep -- the fused sm_100a launch
sequence -- instead of ~hundreds of eager ATen/cuDNN launches and five .item() syncs per step: the
five loss scalars stay on the device and are read back once per epoch. Additive flags: --synthetic N
(train on N synthetic pairs, no dataset needed) and --image_size."""
import argparse
import json
import os
import time

import numpy as np
import torch
from torch.optim import lr_scheduler
from torch.utils.data import DataLoader, Dataset

from .discriminators.discriminators import create_disc
from .generators.generators import create_gen
from .optim import FusedAdam
from .step import TrainStep
from .util import init_weights, mkdir

opt = None  # module-global read by Train_GAN.get_scheduler, like the reference (train.py:191-195)


# ---------------------------------------------------------------------------------- data parallel (one rank per GPU)
def dp_env():
    """(rank, world, local_rank) from the torchrun environment; (0, 1, 0) for a plain `python -m` launch."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def global_real_label(world, batch, h5, w5, generator=None):
    """GANLoss.get_target_tensor (generators/generators.py:52-63) for the GLOBAL batch: one CPU-RNG draw of
    (world*batch, 1, h5, w5), identical on every rank (same seed); rank r trains on rows [r*batch, (r+1)*batch)."""
    return torch.clamp(torch.normal(1.0, 0.02, size=(world * batch, 1, h5, w5), generator=generator), 0, 1)


def rank_rows(t, rank, batch):
    return t[rank * batch:(rank + 1) * batch].contiguous()


class SyntheticPairs(Dataset):
    """Synthetic (source, target) pairs in the dataset's value ranges (PairedDataset.py:52-58,86)."""

    def __init__(self, n, size, in_nc=3, out_nc=3, seed=21):
        self.n, self.size, self.in_nc, self.out_nc, self.seed = n, size, in_nc, out_nc, seed

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed + i)
        return (torch.rand(self.in_nc, self.size, self.size, generator=g) * 2 - 1,
                torch.rand(self.out_nc, self.size, self.size, generator=g))


class Train_GAN:
    """GAN model for the Pix2Pix-style training (reference train.py:22-227)."""

    def __init__(self, opt_, traindataset):
        global opt
        if opt is None:
            opt = opt_
        o = opt_
        self.opt = o
        self.rank, self.world, _ = dp_env()
        self.world = self.world if (torch.distributed.is_available() and torch.distributed.is_initialized()) else 1
        if len(traindataset) == 0:
            raise ValueError("empty training set")
        # o.batch_size is the PER-RANK batch (weak scaling: global batch = world * batch_size). Like the reference
        # (train.py:29) the last partial batch of an epoch is kept; under data parallelism every rank must run the same
        # number of equally sized steps, so the sampler pads to a multiple of world * batch_size instead.
        self.sampler = None
        if self.world > 1:
            from torch.utils.data.distributed import DistributedSampler
            self.sampler = DistributedSampler(traindataset, num_replicas=self.world, rank=self.rank, shuffle=True,
                                              seed=21, drop_last=False)
        self.dataset = DataLoader(dataset=traindataset, batch_size=o.batch_size, shuffle=self.sampler is None,
                                  sampler=self.sampler, num_workers=o.threads, drop_last=self.world > 1,
                                  pin_memory=True,
                                  persistent_workers=o.threads > 0)   # workers survive the epoch: 0.25 s/epoch at 8 threads
        if self.world > 1 and len(self.dataset) == 0:
            raise ValueError(f"{len(traindataset)} samples cannot fill one batch of {o.batch_size} on each of "
                             f"{self.world} ranks")
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.activation = o.loss == "ls"                       # reference train.py:33
        self.return_filter = o.version == 2
        self.netG = create_gen(o.gen, o.input_dim, o.output_dim, o.nf, self.activation).to(self.device)
        init_weights(self.netG)
        self.netD = create_disc("patch", o.input_dim, o.output_dim, o.nf, return_filter=self.return_filter,
                                activation=self.activation).to(self.device)
        init_weights(self.netD)
        if self.world > 1:                                       # every replica starts from rank 0's weights
            for net in (self.netG, self.netD):
                for p in net.parameters():
                    torch.distributed.broadcast(p.data, 0)
        self.optimizer_G = FusedAdam(self.netG, lr=o.lr, betas=(o.beta1, 0.99))
        self.optimizer_D = FusedAdam(self.netD, lr=o.lr, betas=(o.beta1, 0.99))
        self.optimizers = [self.optimizer_G, self.optimizer_D]
        self.schedulers = [self.get_scheduler(x) for x in self.optimizers]
        self.gen_loss, self.disc_loss, self.l1_loss, self.per_loss, self.gp_loss = [], [], [], [], []
        self.step = None
        self.steps_by_batch = {}                               # per-rank batch size -> TrainStep (full batch + ragged tail)
        self.aug_rng = torch.Generator().manual_seed(21 + self.rank)   # util.py:8-11 seeds everything with 21
        self.label_rng = torch.Generator().manual_seed(21)
        self._warned_tail = False
        if o.continue_training:
            ckpt = torch.load(os.path.join(f"{o.data.rsplit('/', 1)[0]}/models", o.folder_load, "final_model.pth"),
                              map_location=self.device)
            self.netG.load_state_dict(ckpt["gen"])
            self.netD.load_state_dict(ckpt["disc"])
            self._pending_opt_state = (ckpt["optimizerG_state_dict"], ckpt["optimizerD_state_dict"])
        else:
            self._pending_opt_state = None

    def _build_step(self, h, w, n=None):
        """TrainStep for a per-rank batch of n (default: the full batch). The ragged last batch of an epoch gets its
        own engines; parameters, Adam state and gradient arenas are shared through the modules' ParamStore."""
        o = self.opt
        n = o.batch_size if n is None else n
        if (n, h, w) in self.steps_by_batch:
            self.step = self.steps_by_batch[(n, h, w)]
            return self.step
        vgg_blocks = None
        if o.version != 2 and o.lambda_per != 0:
            # train.py:48-49: VGGPerceptualLoss(resize=True) -- frozen VGG16 slices (ImageNet weights when available)
            if not hasattr(self, "perceptual"):
                from .util import VGGPerceptualLoss
                self.perceptual = VGGPerceptualLoss(resize=True)
            vgg_blocks = self.perceptual.blocks
        step = TrainStep(self.netG, self.netD, n, h, w, loss=o.loss, version=o.version,
                         lambda_a=o.lambda_a, lambda_gp=o.lambda_gp, lambda_per=o.lambda_per, w_per=o.w_per,
                         lr=o.lr, beta1=o.beta1, label_smoothing=not o.no_label_smoothing,
                         vgg_blocks=vgg_blocks)
        if step.world > 1 and step.label_smoothing and o.loss in ("ls", "ce"):
            # the smoothed real-label tensor is ONE draw for the global batch; each rank takes its rows
            lab = global_real_label(step.world, n, step.h5, step.w5, generator=self.label_rng)
            step.set_label(rank_rows(lab, self.rank, n))
        self.steps_by_batch[(n, h, w)] = step
        self.step = step
        if self._pending_opt_state is not None:
            self.optimizer_G.load_state_dict(self._pending_opt_state[0])
            self.optimizer_D.load_state_dict(self._pending_opt_state[1])
            self._pending_opt_state = None
        return step

    def _step_for(self, n, h, w):
        """The TrainStep serving a batch of n images, or None when the batch has to be skipped: the reference caches
        its smoothed real-label tensor at the FIRST batch's shape (generators/generators.py:54-61), so with label
        smoothing and ls / ce a smaller last batch makes its `expand_as` raise; here that batch is dropped with a
        warning instead. Without smoothing (or with w / hinge) the partial batch trains, as in the reference."""
        o = self.opt
        if n != o.batch_size and self.step is not None and not o.no_label_smoothing and o.loss in ("ls", "ce"):
            if not self._warned_tail:
                print(f"\twarning: last batch of {n} < batch_size {o.batch_size} skipped (the reference's cached "
                      f"smoothed label tensor cannot serve it; use --no_label_smoothing to train on it)")
                self._warned_tail = True
            return None
        return self._build_step(h, w, n)

    def _alpha(self, n):
        """GP alpha (util.py:79): one draw per global batch on the CUDA generator, identical on every rank; this
        rank's rows. Single process: None (TrainStep draws it, same call as the reference)."""
        if self.world == 1:
            return None
        return rank_rows(torch.rand(self.world * n, 1, device=self.device), self.rank, n)

    def train(self, o):
        for i in range(o.total_epochs):
            epoch = i + o.initial_epoch
            t1 = time.time()
            if self.rank == 0:
                print("==training epoch ", epoch)
            regularize = (o.reg_every != 0) and (epoch % o.reg_every == 0) and (o.lambda_gp != 0)
            accum, steps = None, 0
            if self.sampler is not None:
                self.sampler.set_epoch(epoch)
            it = iter(self.dataset)
            nxt = next(it, None)
            while nxt is not None:
                # one-batch lookahead: the next batch's H2D copy is issued before this step's kernels and runs under them
                batch, nxt = nxt, next(it, None)
                host_path = batch[0].dtype == torch.float32 and batch[1].dtype == torch.float32
                if host_path:
                    a, b = batch[0].contiguous(), batch[1].contiguous()
                    step = self._step_for(a.shape[0], a.shape[2], a.shape[3])
                    if step is None:
                        continue
                    if accum is None:
                        accum = torch.zeros_like(step.losses)
                    step.lr = self.optimizer_G.param_groups[0]['lr']
                    ahead = None
                    if nxt is not None and nxt[0].dtype == torch.float32 and nxt[0].is_contiguous() \
                            and nxt[1].dtype == torch.float32 and nxt[1].is_contiguous() \
                            and nxt[0].shape == a.shape:
                        ahead = (nxt[0], nxt[1])
                    accum += step.step_from_host(a, b, regularize=regularize, prefetch=ahead, read=False,
                                                 alpha=self._alpha(a.shape[0]))
                    steps += 1
                    continue
                if batch[0].dtype == torch.uint8:
                    # raw uint8 HWC pair from datasets.PairedDataset(raw=True): ToTensor / Normalize and -- with
                    # augmentation on -- HorizontalFlip + Affine run on the device (augment.py; PairedDataset.py:80-92)
                    from .augment import augment_pair, identity_params, sample_params
                    n, h, w = batch[0].shape[:3]
                    aug = getattr(self.dataset.dataset, "aug", False)
                    q = sample_params(n, h, w, generator=self.aug_rng) if aug else identity_params(n)
                    real_A, real_B = augment_pair(batch[0].to(self.device, non_blocking=True),
                                                  batch[1].to(self.device, non_blocking=True), q)
                else:
                    real_A = batch[0].to(self.device, non_blocking=True).float().contiguous()
                    real_B = batch[1].to(self.device, non_blocking=True).float().contiguous()
                step = self._step_for(real_A.shape[0], real_A.shape[2], real_A.shape[3])
                if step is None:
                    continue
                if accum is None:
                    accum = torch.zeros_like(step.losses)
                step.lr = self.optimizer_G.param_groups[0]['lr']
                losses = step.step(real_A, real_B, regularize=regularize, alpha=self._alpha(real_A.shape[0]))
                accum += losses
                steps += 1
            for scheduler in self.schedulers:
                scheduler.step()
            lr = self.optimizers[0].param_groups[0]['lr']
            if steps == 0:
                raise RuntimeError("the epoch ran zero training steps (every batch was skipped)")
            mean = accum / steps
            if self.world > 1:                                   # the logged scalars are reduced once per epoch
                torch.distributed.all_reduce(mean)
                mean /= self.world
            m = mean.tolist()                                    # the epoch's only sync
            diff = time.time() - t1
            if self.rank == 0:
                print(f"\tloss functions => D:{m[0]:.5f}, G:{m[2]:.5f}, L1:{m[3]:.5f}, gp:{m[1]:.5f}, per:{m[4]:.5f}")
                print(f'\tlearing rate: {lr:.5f}')
                print(f"\ttook {diff:.2f} seconds")
                print(f"\tapproximately {diff * (o.total_epochs - epoch):.2f} seconds left")
            self.gen_loss.append(m[2])
            self.disc_loss.append(m[0])
            self.l1_loss.append(m[3])
            self.per_loss.append(m[4])
            self.gp_loss.append(m[1])
            if self.rank == 0 and o.checkpoint_interval != -1 and epoch % o.checkpoint_interval == 0:
                self.save_model(f"{o.data.rsplit('/', 1)[0]}/checkpoints/{o.folder_save}/model_{epoch}.pth")

    @staticmethod
    def get_scheduler(optimizer):
        milestone = np.int16(np.linspace(opt.epoch_constant, opt.total_epochs, 11)[:-1])
        return lr_scheduler.MultiStepLR(optimizer, milestones=list(milestone), gamma=0.8)

    def save_model(self, modelpath):
        if not os.path.exists(modelpath.rsplit('/', 1)[0]):
            mkdir(modelpath.rsplit('/', 1)[0])
        torch.save({'gen': self.netG.state_dict(), 'disc': self.netD.state_dict(),
                    'optimizerG_state_dict': self.optimizer_G.state_dict(),
                    'optimizerD_state_dict': self.optimizer_D.state_dict()}, modelpath)

    def save_arrays(self, path):
        for name, arr in (("genloss", self.gen_loss), ("discloss", self.disc_loss), ("l1loss", self.l1_loss),
                          ("perloss", self.per_loss), ("gploss", self.gp_loss)):
            np.save(os.path.join(path, name), np.asarray(arr))

    def save_hyper_params(self, folderpath, o):
        with open(os.path.join(folderpath, 'params.txt'), 'w') as file:
            file.write(json.dumps(o.__dict__))


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument("--data", default="./data", help="dataset directory")
    parser.add_argument("--batch_size", type=int, default=4, help="training batch size")
    parser.add_argument("--input_dim", type=int, default=3, help="input depth size")
    parser.add_argument("--output_dim", type=int, default=3, help="output depth size")
    parser.add_argument("--initial_epoch", type=int, default=1, help="starting epoch")
    parser.add_argument("--total_epochs", type=int, default=135, help="total epochs we're training for")
    parser.add_argument("--epoch_constant", type=int, default=25, help="epochs with constant learning rate")
    parser.add_argument("--lr", type=float, default=0.001, help="learning rate")
    parser.add_argument("--no_label_smoothing", default=False, action='store_true', help="no one sided label smoothing")
    parser.add_argument("--beta1", type=float, default=0.9, help="beta1 for our Adam optimizer")
    parser.add_argument("--threads", type=int, default=8, help="cpu threads for loading the dataset")
    parser.add_argument("--lambda_a", type=float, default=1, help="L1 loss coefficient")
    parser.add_argument('--lambda_gp', type=float, default=0.01, help="gradient penalty coefficient")
    parser.add_argument("--lambda_per", type=float, default=1, help="perceptual loss coefficient")
    parser.add_argument('--w_per', nargs=4, type=float, default=[0, .1, .3, .6], help='perceptual weights')
    parser.add_argument("--gen", default="UNet++", choices=["UNet++", "UNet", "BCDUNet"], help="generator architecture")
    parser.add_argument("--nf", type=int, default=64, help="base number of filters")
    parser.add_argument("--loss", default="ls", choices=["ls", "ce", "w", "hinge"], help="loss function for ganloss")
    parser.add_argument("--no_aug", default=False, action='store_true', help="do not augment the dataset")
    parser.add_argument("--target", default="rgb", choices=["ch", "rgb"], help="target image format")
    parser.add_argument("-v", "--version", type=int, default=1, choices=[1, 2], help="version of the tactile GAN")
    parser.add_argument("--folder_save", default="pix2obj", help="where we want to save the model to")
    parser.add_argument("--folder_load", default="pix2obj", help="where we want to load the model from")
    parser.add_argument("--checkpoint_interval", type=int, default=-1, help="interval between model checkpoints")
    parser.add_argument("--continue_training", default=False, action='store_true', help="load weights before training")
    parser.add_argument('--reg_every', type=int, default=1, help='how frequently we regularize using gp')
    # additive (not in the reference)
    parser.add_argument("--synthetic", type=int, default=0, help="train on N synthetic pairs instead of --data")
    parser.add_argument("--image_size", type=int, default=256, help="side of the synthetic pairs")
    return parser


def replicas_agree(nets, device):
    """Data-parallel safety net: every rank must hold bit-identical weights (same reduced gradients, same Adam
    kernel). One 2-element all-reduce at the end of training."""
    chk = torch.zeros(2, device=device, dtype=torch.float64)
    for net in nets:
        for p in net.parameters():
            chk[0] += p.detach().double().sum()
            chk[1] += p.detach().double().abs().sum()
    lo, hi = chk.clone(), chk.clone()
    torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN)
    torch.distributed.all_reduce(hi, op=torch.distributed.ReduceOp.MAX)
    return bool((lo == hi).all())


def main(argv=None):
    """`python -m tactile_gan_b200.train <flags>` on one GPU, or one rank per GPU:
    `python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 -m tactile_gan_b200.train <flags>`
    (--batch_size is per rank; NCCL gradient all-reduce inside TrainStep; rank 0 writes the files)."""
    global opt
    opt = build_parser().parse_args(argv)
    rank, world, local = dp_env()
    if world > 1:
        torch.cuda.set_device(local)
        if not torch.distributed.is_initialized():
            torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    if opt.synthetic > 0:
        train_set = SyntheticPairs(opt.synthetic, opt.image_size, opt.input_dim, opt.output_dim)
    else:
        from .datasets.datasets import get_dataset
        train_set = get_dataset(os.path.join(opt.data, "train", "source"), opt, mode='train', raw=True)
    experiment = Train_GAN(opt, train_set)
    root = opt.data.rsplit('/', 1)[0]
    save_path = os.path.join(f"{root}/models", opt.folder_save)
    if rank == 0:
        mkdir(os.path.join(f"{root}/checkpoints", opt.folder_save))
        mkdir(save_path)
    experiment.train(opt)
    if world > 1 and not replicas_agree([experiment.netG, experiment.netD], experiment.device):
        raise RuntimeError("data-parallel replicas diverged: ranks hold different weights after training")
    if rank == 0:
        experiment.save_model(os.path.join(save_path, "final_model.pth"))
        experiment.save_arrays(save_path)
        experiment.save_hyper_params(save_path, opt)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
