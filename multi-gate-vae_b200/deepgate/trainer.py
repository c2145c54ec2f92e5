"""Trainer with the reference's interface (trainer.py:20-278): ``Trainer(args, model, ...)``,
``set_training_args``, ``run_batch``, ``train``, ``save`` / ``load`` / ``resume``.

Differences, all deliberate:
  * gradients ARE synchronised across ranks: one flat-buffer NCCL all-reduce (sum / world) per step
    -- the reference shards the data per rank but never all-reduces;
  * the edge "split" is an O(E) permutation instead of the N x N mask of preprocessing.py:56-69;
  * the func loss (and, for VAE models, reparam + KL) runs in the fused CUDA kernel;
  * replicas start from rank 0's parameters and buffers (broadcast at construction and after ``load``), as DDP does --
    the reference sets no seed (train.py) so its ranks would train different models;
  * a batch that fails the asynchronous input validation (node id / level / edge end out of range) does not update the
    parameters: the device-side flag is handed to the fused Adam kernel as its ``found_inf`` operand (no host sync), and
    the training loop raises when it next reads results (the reference raises an IndexError before the step).
"""
import os
import time

import numpy as np
import torch
from torch import nn

from . import ops
from .data import CudaPrefetcher, DataLoader, DeferredScalars
from .utils.utils import AverageMeter


def split_edges(batch):
    """``general_train_test_split_edges`` with val_ratio = test_ratio = 0 (preprocessing.py:8-83):
    every edge is a training edge, in random order."""
    ei = batch.edge_index
    if ei.is_cuda:
        batch.train_pos_edge_index = ops.permute_edges(ei)       # one launch, no sort (csrc/recon.cu)
    else:                                                        # host-side callers (data preparation, CPU tests of the loop)
        batch.train_pos_edge_index = ei[:, torch.randperm(ei.size(1))]
    return batch


class _WeightedSum(torch.autograd.Function):
    """sum_i w_i x_i of 0-dim losses: stack + dot forward, one scale backward (the expression w0 * a + w1 * b + w2 * c is five
    tiny launches forward and eight backward)."""

    @staticmethod
    def forward(ctx, wvec, *xs):
        ctx.save_for_backward(wvec)
        return torch.dot(torch.stack([x.reshape(()) for x in xs]), wvec)

    @staticmethod
    def backward(ctx, g):
        (wvec,) = ctx.saved_tensors
        return (None,) + tuple((g * wvec).unbind(0))


class FlatGradAllReduce(object):
    """Mean of the gradients over ranks through one flat fp32 buffer (0.96-1.56 MB for these models)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.flat = None

    def __call__(self):
        if not (torch.distributed.is_available() and torch.distributed.is_initialized()):
            return
        world = torch.distributed.get_world_size()
        if world == 1:
            return
        # one cat, one all-reduce, one scale, one multi-tensor copy back (instead of two small ops per parameter)
        for p in self.params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        grads = [p.grad for p in self.params]
        self.flat = torch.cat([g.reshape(-1) for g in grads])
        if torch.distributed.get_backend() == "nccl":
            torch.distributed.all_reduce(self.flat, op=torch.distributed.ReduceOp.AVG)      # the mean inside the collective
        else:
            torch.distributed.all_reduce(self.flat, op=torch.distributed.ReduceOp.SUM)
            self.flat.div_(world)
        views, off = [], 0
        for g in grads:
            n = g.numel()
            views.append(self.flat[off:off + n].view_as(g))
            off += n
        torch._foreach_copy_(grads, views)


class Logger(object):
    def __init__(self, path):
        self.path = path

    def write(self, txt):
        with open(self.path, "a") as f:
            f.write(txt)


class Trainer(object):
    def __init__(self, args, model, training_id="default", save_dir="./exp", lr=1e-4,
                 rc_prob_func_weight=[1.0, 4.0, 2.0], emb_dim=128, device="cpu", batch_size=32, num_workers=0,
                 distributed=True):
        self.args = args
        self.emb_dim = emb_dim
        self.device = device
        self.lr = lr
        self.lr_step = -1
        self.rc_prob_func_weight = list(rc_prob_func_weight)
        self.log_dir = os.path.join(save_dir, training_id)
        os.makedirs(self.log_dir, exist_ok=True)
        self.log_path = os.path.join(self.log_dir, "log-{}.txt".format(time.strftime("%Y-%m-%d-%H-%M")))
        self.batch_size = batch_size
        self.num_workers = num_workers
        self.distributed = distributed
        self.local_rank, self.rank, self.world_size = 0, 0, 1
        if self.distributed:
            self.local_rank = int(os.environ.get("LOCAL_RANK", 0))
            self.device = "cuda:%d" % self.local_rank
            torch.cuda.set_device(self.local_rank)
            if not torch.distributed.is_initialized():
                torch.distributed.init_process_group(backend="nccl", init_method="env://")
            self.world_size = torch.distributed.get_world_size()
            self.rank = torch.distributed.get_rank()
            print("Training in distributed mode. Device {}, Process {:}, total {:}.".format(
                self.device, self.rank, self.world_size))
        else:
            print("Training in single device: ", self.device)
        self.reg_loss = nn.L1Loss().to(self.device)
        self.clf_loss = nn.BCELoss().to(self.device)
        # the optimizer is created before .to(device), like the reference (trainer.py:73-76)
        self.optimizer = torch.optim.Adam(model.parameters(), lr=self.lr)
        self.model = model.to(self.device)
        if str(self.device).startswith("cuda"):
            for group in self.optimizer.param_groups:        # one multi-tensor kernel per step instead of ~100 small ones
                group["fused"] = True
                group["foreach"] = False
        self.model_epoch = 0
        self.kl_weight = 0.0                 # the reference computes KL for VAE models but leaves it out of the total (trainer.py:167,227-231)
        self.grad_sync = FlatGradAllReduce(self.model.parameters())
        self.sync_replicas()
        if self.local_rank == 0:
            self.logger = Logger(self.log_path)

    def sync_replicas(self):
        """Rank 0's parameters AND buffers (BatchNorm running statistics) to every rank, in one flat broadcast."""
        if not (torch.distributed.is_available() and torch.distributed.is_initialized()):
            return
        if torch.distributed.get_world_size() == 1:
            return
        tensors = [t for t in list(self.model.parameters()) + list(self.model.buffers())]
        with torch.no_grad():
            for dt in sorted({t.dtype for t in tensors}, key=str):
                group = [t for t in tensors if t.dtype == dt]
                flat = torch.cat([t.detach().reshape(-1) for t in group])
                torch.distributed.broadcast(flat, src=0)
                off = 0
                for t in group:
                    t.copy_(flat[off:off + t.numel()].view_as(t))
                    off += t.numel()

    def set_training_args(self, rc_prob_func_weight=[], lr=-1, lr_step=-1, device="null"):
        if len(rc_prob_func_weight) == 3 and list(rc_prob_func_weight) != self.rc_prob_func_weight:
            print("[INFO] Update rc_prob_func_weight from {} to {}".format(self.rc_prob_func_weight, rc_prob_func_weight))
            self.rc_prob_func_weight = list(rc_prob_func_weight)
        if lr > 0 and lr != self.lr:
            print("[INFO] Update learning rate from {} to {}".format(self.lr, lr))
            self.lr = lr
            for group in self.optimizer.param_groups:
                group["lr"] = self.lr
        if lr_step > 0 and lr_step != self.lr_step:
            print("[INFO] Update learning rate step from {} to {}".format(self.lr_step, lr_step))
            self.lr_step = lr_step
        if device != "null" and device != self.device:
            print("[INFO] Update device from {} to {}".format(self.device, device))
            self.device = device
            self.model = self.model.to(self.device)
            self.reg_loss = self.reg_loss.to(self.device)
            self.clf_loss = self.clf_loss.to(self.device)

    # ------------------------------------------------------------------ checkpoints
    def save(self, path):
        torch.save({"epoch": self.model_epoch, "state_dict": self.model.state_dict(),
                    "optimizer": self.optimizer.state_dict()}, path)

    def load(self, path):
        checkpoint = torch.load(path, map_location=lambda storage, loc: storage)
        self.optimizer.load_state_dict(checkpoint["optimizer"])
        for group in self.optimizer.param_groups:
            self.lr = group["lr"]
        self.model_epoch = checkpoint["epoch"]
        self.model.load(path)
        self.sync_replicas()
        print("[INFO] Continue training from epoch {:}".format(self.model_epoch))
        return path

    def resume(self):
        path = os.path.join(self.log_dir, "model_last.pth")
        if not os.path.exists(path):
            return False
        self.load(path)
        return True

    # ------------------------------------------------------------------ one batch
    def run_batch(self, batch, neg_edge_index=None):
        if getattr(batch, "train_pos_edge_index", None) is None:
            batch = split_edges(batch)
        hs, hf = self.model(batch)
        loss, pred_bin, gt_bin = self.model.recon_loss(hs, batch.train_pos_edge_index, neg_edge_index)
        if hasattr(self.model, "pred_prob_loss"):                  # fused MLP + clamp + L1 (csrc/readout.cu)
            prob, prob_loss = self.model.pred_prob_loss(hf, batch["prob"])
        else:
            prob = self.model.pred_prob(hf)
            prob_loss = self.reg_loss(prob, batch["prob"])
        # func loss: 1 - cos -> z-norm -> L1 against z-norm(tt_sim)   (trainer.py:157-163), fused kernel
        _, _, _, func_loss = ops.vae_func_loss(hf=hf, tt_pair_index=batch["tt_pair_index"], tt_sim=batch["tt_sim"])
        status = {"recon_loss": loss, "pred_bin": pred_bin, "gt_bin": gt_bin, "prob_loss": prob_loss,
                  "func_loss": func_loss, "hs": hs, "hf": hf}
        # Variational branch (trainer.py:145-151): models that stash s_mu / s_logstd / t_mu / t_logstd (DirectedGVAE.sample,
        # digvae_model.py:134-142; here the level models built with ``variational=True``) report the KL term, which the
        # fused reparam + KL kernel produced during forward.
        if self._is_vae():
            status["kl_loss"] = self.model.kl_loss()
        return status

    def _is_vae(self):
        name = getattr(self.args, "model", "") if self.args is not None else ""
        return ("VAE" in (name or "") or getattr(self.model, "variational", False)) and hasattr(self.model, "kl_loss")

    def total_loss(self, status):
        w = self.rc_prob_func_weight
        terms = [status["recon_loss"], status["prob_loss"], status["func_loss"]]
        weights = [float(w[0]), float(w[1]), float(w[2])]
        if self.kl_weight and "kl_loss" in status:
            terms.append(status["kl_loss"])
            weights.append(float(self.kl_weight))
        if all(t.is_cuda and t.dtype == torch.float32 for t in terms):
            key = (tuple(weights), terms[0].device)
            if getattr(self, "_wvec_key", None) != key:
                self._wvec_key, self._wvec = key, torch.tensor(weights, dtype=torch.float32, device=terms[0].device)
            return _WeightedSum.apply(self._wvec, *terms)
        total = weights[0] * terms[0]
        for wi, t in zip(weights[1:], terms[1:]):
            total = total + wi * t
        return total

    def train_step(self, batch, neg_edge_index=None):
        """zero_grad -> run_batch -> backward -> gradient all-reduce -> Adam step.  Returns the status dict."""
        self.optimizer.zero_grad()
        status = self.run_batch(batch, neg_edge_index)
        loss = self.total_loss(status)
        loss.backward()
        self.grad_sync()
        self._guarded_step()
        status["loss"] = loss.detach()
        return status

    def _guarded_step(self):
        """optimizer.step() that leaves the parameters untouched when this step's batch failed the deferred input validation:
        the flag word goes to the fused Adam kernel as ``found_inf`` (the GradScaler hook), so nothing synchronises."""
        from .schedule import error_word
        fused = str(self.device).startswith("cuda") and all(g.get("fused") for g in self.optimizer.param_groups)
        if fused:
            if getattr(self, "_found_inf", None) is None:
                self._found_inf = torch.zeros((), dtype=torch.float32, device=self.device)
            torch.ne(error_word(torch.device(self.device))[0], 0, out=self._found_inf)      # one tiny launch, no sync
            self.optimizer.found_inf = self._found_inf
            self.optimizer.grad_scale = None
            try:
                self.optimizer.step()
            finally:
                del self.optimizer.found_inf, self.optimizer.grad_scale
        else:
            self.optimizer.step()

    # ------------------------------------------------------------------ loop
    def train(self, num_epoch, train_dataset, val_dataset):
        def loader(ds):
            if self.distributed and self.world_size > 1:
                sampler = torch.utils.data.distributed.DistributedSampler(ds, num_replicas=self.world_size, rank=self.rank)
                return DataLoader(ds, batch_size=self.batch_size, shuffle=False, drop_last=True,
                                  num_workers=self.num_workers, sampler=sampler)
            return DataLoader(ds, batch_size=self.batch_size, shuffle=True, drop_last=True, num_workers=self.num_workers)

        from .schedule import check_deferred_errors, error_word
        loaders = {"train": loader(train_dataset), "val": loader(val_dataset)}
        meters = {k: AverageMeter() for k in ("time", "recon", "prob", "func", "acc")}
        on_cuda = str(self.device).startswith("cuda")

        def account(vals):
            # two steps late (DeferredScalars, lag 2): the losses, the accuracy and the deferred input-validation word of a step are read
            # while the next step is already queued -- the reference's .item() / .cpu() per batch (trainer.py:219-226) drains the
            # device once per step.  A batch that failed validation never reached the parameters (_guarded_step).
            recon, prob, func, acc = vals[:4]
            if len(vals) > 4 and vals[4]:
                check_deferred_errors()
            meters["recon"].update(recon)
            meters["prob"].update(prob)
            meters["func"].update(func)
            meters["acc"].update(acc)

        print("[INFO] Start training, lr = {:.4f}".format(self.optimizer.param_groups[0]["lr"]))
        for epoch in range(num_epoch):
            for phase in ("train", "val"):
                self.model.train(phase == "train")
                reader = DeferredScalars(self.device, lag=2)
                t0 = time.time()
                # host -> device copy of batch i + 1 behind the step on batch i (the reference: blocking batch.to(device), trainer.py:201)
                for batch in CudaPrefetcher(loaders[phase], self.device):
                    if phase == "train":
                        status = self.train_step(batch)
                    else:
                        with torch.no_grad():
                            status = self.run_batch(batch)
                    acc = (status["pred_bin"] == status["gt_bin"]).to(torch.float32).mean()
                    vals = [status["recon_loss"], status["prob_loss"], status["func_loss"], acc]
                    if on_cuda:
                        vals.append(error_word(self.device)[0])
                    done = reader.push(*vals)
                    if done is not None:
                        account(done)
                    meters["time"].update(time.time() - t0)      # wall time per batch, everything included
                    t0 = time.time()
                for done in reader.drain():
                    account(done)
                check_deferred_errors()
                if phase == "train" and self.model_epoch % 10 == 0 and self.rank == 0:
                    self.save(os.path.join(self.log_dir, "model_{:}.pth".format(self.model_epoch)))
                    self.save(os.path.join(self.log_dir, "model_last.pth"))
                if self.local_rank == 0:
                    self.logger.write("{}| Epoch: {:}/{:} |Recon: {:.4f} |ACC: {:.2f} |Prob: {:.4f} |Func: {:.4f}|Net: {:.2f}s\n".format(
                        phase, epoch, num_epoch, meters["recon"].avg, meters["acc"].avg * 100, meters["prob"].avg,
                        meters["func"].avg, meters["time"].avg))
            self.model_epoch += 1
            if self.lr_step > 0 and self.model_epoch % self.lr_step == 0:
                self.lr *= 0.1
                for group in self.optimizer.param_groups:
                    group["lr"] = self.lr
