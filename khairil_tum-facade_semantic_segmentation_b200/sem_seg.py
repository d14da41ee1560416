"""The SSG semantic-segmentation network that calls the hot path.

Architecture and attribute names follow /root/reference/models/pointnet2_sem_seg.py:6-50
(4 set-abstraction levels, 4 feature-propagation levels, Conv1d/BN/Dropout/Conv1d head,
log_softmax), so its ``state_dict`` is interchangeable with the reference's ``get_model``.
The reference file itself also runs unchanged on top of ``models/pointnet2_utils.py``; this
copy exists because the reference tree is not present on the GPU box.  The head and the loss
are ordinary PyTorch (they are outside the hot path).
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .modules import PointNetFeaturePropagation, PointNetSetAbstraction

# (npoint, radius, nsample, mlp) -- pointnet2_sem_seg.py:9-12
SA_LEVELS = ((1024, 0.1, 32, (32, 32, 64)), (256, 0.2, 32, (64, 64, 128)),
             (64, 0.4, 32, (128, 128, 256)), (16, 0.8, 32, (256, 256, 512)))
# (in_channel, mlp) for fp4, fp3, fp2, fp1 -- pointnet2_sem_seg.py:13-16
FP_LEVELS = ((768, (256, 256)), (384, (256, 256)), (320, (256, 128)), (128, (128, 128, 128)))


_GEO_STREAMS = {}


def _geometry_stream(dev):
    s = _GEO_STREAMS.get(dev.index)
    if s is None:
        s = _GEO_STREAMS[dev.index] = torch.cuda.Stream(device=dev)
    return s


_GEO_HELPERS = {}


def _geometry_helpers(dev):
    """two helper streams per device for get_model.geometry_all(fork=True)"""
    h = _GEO_HELPERS.get(dev.index)
    if h is None:
        h = _GEO_HELPERS[dev.index] = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
    return h


class _NoWait:
    @staticmethod
    def wait_event(_):
        pass


class get_model(nn.Module):
    def __init__(self, num_classes, num_extra_features, sa_cls=PointNetSetAbstraction,
                 fp_cls=PointNetFeaturePropagation):
        super().__init__()
        width = 6 + 3 + num_extra_features
        for level, (npoint, radius, nsample, mlp) in enumerate(SA_LEVELS, start=1):
            self.add_module("sa%d" % level, sa_cls(npoint, radius, nsample, width, list(mlp), False))
            width = mlp[-1] + 3
        for level, (cin, mlp) in zip((4, 3, 2, 1), FP_LEVELS):
            self.add_module("fp%d" % level, fp_cls(cin, list(mlp)))
        self.conv1 = nn.Conv1d(128, 128, 1)
        self.bn1 = nn.BatchNorm1d(128)
        self.drop1 = nn.Dropout(0.5)
        self.conv2 = nn.Conv1d(128, num_classes, 1)

    overlap_geometry = True     # run the coordinate-only index pipeline on a side stream, ahead of the MLPs

    def _geometry_ahead(self, xyz0):
        """FPS -> ball query for the four levels, then the four 3-NN searches, on a side stream: all of it depends
        on coordinates only, so it runs concurrently with the feature MLPs, which wait per level."""
        dev = xyz0.device
        main = torch.cuda.current_stream(dev)
        side = _geometry_stream(dev)
        side.wait_stream(main)           # inputs ready; also orders reuse of last step's index buffers
        geo, nn3, events = [], [], []
        with torch.cuda.stream(side):
            coords = [xyz0]
            for sa in (self.sa1, self.sa2, self.sa3, self.sa4):
                g = sa.geometry(coords[-1])
                geo.append(g)
                coords.append(g[0].permute(0, 2, 1))
                events.append(side.record_event())
            for fine, coarse in ((3, 4), (2, 3), (1, 2), (0, 1)):
                nn3.append(PointNetFeaturePropagation.neighbours(coords[fine], coords[coarse]))
                events.append(side.record_event())
        return geo, nn3, events, main

    fork_geometry = os.environ.get("PN2_FORK_GEOMETRY", "1") != "0"      # geometry_all: ball queries and 3-NN searches on helper streams beside the FPS chain

    def geometry_all(self, xyz0, fork=None):
        """The whole coordinate-only index pipeline of one batch, issued from the CURRENT stream: (geo, nn3) as forward()
        takes them through `geometry=`.  A pipelined caller (trainer.SemSegTrainer / SemSegPredictor with pipeline=True)
        runs it for batch i+1 while the feature path of batch i is still busy -- the 1 360-iteration FPS chain keeps
        only B of the 148 SMs busy, so it hides behind the MLPs of the previous batch.

        fork (default: self.fork_geometry): only the four FPS calls depend on each other (level l+1 samples level l's
        centroids); the ball query of level l and the 3-NN search between levels l and l+1 need FPS l only.  They are
        issued on two helper streams that fork after each FPS and join at the end, so the branch takes
        FPS1 + max(ball queries, 3-NN searches) ~ 0.58 ms instead of the 0.97 ms sum -- which had become the length of
        the pipelined inference forward.  Same kernels, same results."""
        fork = self.fork_geometry if fork is None else fork
        sas = (self.sa1, self.sa2, self.sa3, self.sa4)
        if not (fork and xyz0.is_cuda):
            geo, nn3 = [], []
            coords = [xyz0]
            for sa in sas:
                g = sa.geometry(coords[-1])
                geo.append(g)
                coords.append(g[0].permute(0, 2, 1))
            for fine, coarse in ((3, 4), (2, 3), (1, 2), (0, 1)):
                nn3.append(PointNetFeaturePropagation.neighbours(coords[fine], coords[coarse]))
            return geo, nn3
        dev = xyz0.device
        cur = torch.cuda.current_stream(dev)
        balls, knn = _geometry_helpers(dev)
        geo, nn3 = [None] * 4, [None] * 4
        coords = [xyz0]
        for level, sa in enumerate(sas):
            new_xyz = sa.sample(coords[-1])                      # FPS on the issuing stream: the dependent chain
            coords.append(new_xyz.permute(0, 2, 1))
            ready = cur.record_event()
            balls.wait_event(ready)
            with torch.cuda.stream(balls):
                geo[level] = (new_xyz, sa.ball_query(coords[level], new_xyz))
            knn.wait_event(ready)
            with torch.cuda.stream(knn):                         # nn3 is ordered fp4 .. fp1: (fine, coarse) = (level, level + 1)
                nn3[3 - level] = PointNetFeaturePropagation.neighbours(coords[level], coords[level + 1])
        cur.wait_stream(balls)
        cur.wait_stream(knn)
        return geo, nn3

    @staticmethod
    def geometry_tensors(geometry):
        """Flat list of the tensors inside a geometry_all() result (fixed order)."""
        geo, nn3 = geometry
        return [t for g in geo for t in g] + [t for n in nn3 for t in n]

    def forward(self, xyz, geometry=None):
        return self._forward(xyz, geometry, None)

    def forward_loss(self, xyz, target, weight=None, geometry=None):
        """forward() followed by get_loss (pointnet2_sem_seg.py:47-48: F.nll_loss(pred, target, weight=weight)) ->
        (loss, pred, l4_points).  With the fused head (bf16 rows) the loss and its gradient are evaluated inside the head
        kernels (csrc/head.cu): no [B*N, classes] gradient tensor, none of the ~20 small loss kernels; `pred` is then
        detached.  Otherwise the same value through get_loss on the ordinary forward."""
        out = self._forward(xyz, geometry, (target, weight))
        if len(out) == 3:
            return out[2], out[0], out[1]
        pred, l4_points = out
        loss = get_loss()(pred.contiguous().view(-1, pred.shape[-1]), target.view(-1), l4_points, weight)
        return loss, pred, l4_points

    def training_chains(self):
        """The conv chains forward() will run as MLPs, for modules.prepack_mlps (one weight-pack launch per step)."""
        chains = [m.mlp_convs for m in (self.sa1, self.sa2, self.sa3, self.sa4, self.fp4, self.fp3, self.fp2)]
        fp1 = self.fp1
        fused = self.fused_head and type(fp1) is PointNetFeaturePropagation and fp1.head_applies(
            self.conv1, self.bn1, self.conv2, self.conv1.weight)
        chains.append(list(fp1.mlp_convs) + [self.conv1] if fused else fp1.mlp_convs)
        return chains

    def forward_labels(self, xyz, geometry=None):
        """forward() plus the arg-max labels the test loop takes from it (localfunctions.py:398-400:
        `seg_pred.contiguous().cpu().data.max(2)[1]`) -> (labels [B, N] int64, pred, l4_points).  With the fused head the
        labels come out of the head kernel itself (no separate arg-max pass over the [B, N, classes] log-probabilities)."""
        if xyz.is_cuda:
            labels = torch.empty(xyz.shape[0], xyz.shape[2], device=xyz.device, dtype=torch.int64)
            out = self._forward(xyz, geometry, None, labels)
            if len(out) == 3:
                return labels, out[0], out[1]
        else:
            out = self._forward(xyz, geometry, None)
        return out[0].argmax(dim=2), out[0], out[1]

    def _forward(self, xyz, geometry, loss_args, labels_out=None):
        feats = [xyz]
        coords = [xyz[:, :3, :]]
        sas = (self.sa1, self.sa2, self.sa3, self.sa4)
        fps = (self.fp4, self.fp3, self.fp2, self.fp1)
        ahead = xyz.is_cuda and (self.overlap_geometry or geometry is not None) and all(
            type(m) is PointNetSetAbstraction and not m.group_all for m in sas) and all(
            type(m) is PointNetFeaturePropagation for m in fps)
        if geometry is not None:
            if not ahead:
                raise ValueError("geometry= needs the stock set-abstraction / feature-propagation modules on CUDA")
            geo, nn3 = geometry          # computed earlier by geometry_all(); the caller ordered the streams
            main, events = _NoWait, [None] * 8
        elif ahead:
            geo, nn3, events, main = self._geometry_ahead(coords[0])
        for i, sa in enumerate(sas):
            with _lib.sm_budget(self.sa_sm_budget.get(i) if self.sa_sm_budget else None):
                if ahead:
                    main.wait_event(events[i])
                    c, f = sa(coords[-1], feats[-1], geometry=geo[i])
                else:
                    c, f = sa(coords[-1], feats[-1])
            coords.append(c)
            hook = self.feature_grad_hooks.get(i + 1) if self.feature_grad_hooks else None
            if hook is not None and f.requires_grad:
                f.register_hook(hook)
            issued = self.feature_grad_hooks.get("sa%d_issued" % (i + 1)) if self.feature_grad_hooks else None
            if issued is not None:
                issued()           # (forward position: that level's kernels have just been issued on the current stream)
            feats.append(f)
        l4_points = feats[4]
        up = feats[4]
        for i, (fp, fine) in enumerate(zip(fps, (3, 2, 1, 0))):
            skip = feats[fine] if fine > 0 else None
            if fine == 0 and self.fused_head and type(fp) is PointNetFeaturePropagation and fp.head_applies(
                    self.conv1, self.bn1, self.conv2, up):
                # the last level and the head as one chain of rows (csrc/head.cu); log-probabilities come out [B,N,classes]
                if ahead:
                    main.wait_event(events[4 + i])
                if loss_args is not None:
                    pred, loss = fp.forward_with_head(coords[0], coords[1], skip, up, self.conv1, self.bn1, self.drop1,
                                                      self.conv2, neighbours=nn3[i] if ahead else None,
                                                      loss_target=loss_args[0], loss_weight=loss_args[1])
                    return pred, l4_points, loss
                if labels_out is not None:
                    pred = fp.forward_with_head(coords[0], coords[1], skip, up, self.conv1, self.bn1, self.drop1, self.conv2,
                                                neighbours=nn3[i] if ahead else None, labels_out=labels_out)
                    return pred, l4_points, labels_out
                pred = fp.forward_with_head(coords[0], coords[1], skip, up, self.conv1, self.bn1, self.drop1, self.conv2,
                                            neighbours=nn3[i] if ahead else None)
                return pred, l4_points
            if ahead:
                main.wait_event(events[4 + i])
                up = fp(coords[fine], coords[fine + 1], skip, up, neighbours=nn3[i])
            else:
                up = fp(coords[fine], coords[fine + 1], skip, up)
            hook = self.feature_grad_hooks.get("fp%d" % (4 - i)) if self.feature_grad_hooks else None
            if hook is not None and up.requires_grad:
                up.register_hook(hook)
        return self._head(up), l4_points

    sa_sm_budget = None          # {set-abstraction index 0..3: SMs}: grid budget of that level's persistent forward kernels
                                 # (pn2_set_sm_budget); the pipelined trainer / predictor set it while they capture, for the
                                 # levels that run next to the FPS chain of the following batch
    feature_grad_hooks = None    # {level 1..4: hook}: registered on that set-abstraction level's output features when they
                                 # carry a gradient (trainer.SemSegTrainer: early all-reduce of finished gradient buckets);
                                 # {"fp4" | "fp3" | "fp2": hook}: on that feature-propagation level's output (its gradient is
                                 # complete when the next finer level's backward is done: the trainer's anchor for the index
                                 # pipeline of the next batch)
    fused_head = True     # bf16 mode: fp1 + head as one chain of rows on this library's kernels
    rows_head = True      # otherwise: the PyTorch head on point-major rows (no layout copies)

    def _head(self, up):
        """pointnet2_sem_seg.py:36-39.  `up` is [B,128,N]; when it comes from PointNetFeaturePropagation it is a
        permuted view of contiguous point-major rows [B,N,128], and a 1x1 Conv1d on it is a plain matrix product on
        those rows: same parameters (the nn.Conv1d / nn.BatchNorm1d modules own them), same arithmetic, but no
        [B,C,N] <-> [B,N,C] copies and the log-probabilities come out as [B,N,classes] directly.  Outside the
        hot path: ordinary PyTorch kernels."""
        if not self.rows_head:
            x = self.drop1(F.relu(self.bn1(self.conv1(up))))
            x = F.log_softmax(self.conv2(x), dim=1)
            return x.permute(0, 2, 1)
        B, C, N = up.shape
        rows = up.permute(0, 2, 1).reshape(B * N, C)
        x = F.linear(rows, self.conv1.weight.view(self.conv1.out_channels, C), self.conv1.bias)
        bn = self.bn1
        use_batch_stats = bn.training or bn.running_mean is None
        momentum = bn.momentum
        if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
            if momentum is None:
                momentum = 1.0 / float(bn.num_batches_tracked)
        with torch.backends.cudnn.flags(enabled=False):     # the native kernels handle [M, C] rows well; cuDNN's 1x1 path does not
            x = F.batch_norm(x, bn.running_mean if not bn.training or bn.track_running_stats else None,
                             bn.running_var if not bn.training or bn.track_running_stats else None,
                             bn.weight, bn.bias, use_batch_stats, 0.0 if momentum is None else momentum, bn.eps)
        x = self.drop1(F.relu(x))
        x = F.linear(x, self.conv2.weight.view(self.conv2.out_channels, -1), self.conv2.bias)
        return F.log_softmax(x, dim=1).view(B, N, -1)


class get_loss(nn.Module):
    """pointnet2_sem_seg.py:44-50: class-weighted NLL on the log-probabilities (mean over the class weights of
    the targets, F.nll_loss semantics).  On CUDA it is evaluated as gather / weighted sums -- the same value and
    gradient as F.nll_loss(pred, target, weight=weight), whose single-block reduction kernels take ~0.2 ms on a
    32 x 4096-point batch; targets equal to the default ignore_index (-100) contribute nothing, as there."""

    def forward(self, pred, target, trans_feat, weight):
        if not pred.is_cuda or pred.dim() != 2:
            return F.nll_loss(pred, target, weight=weight)
        keep = target != -100
        safe = torch.where(keep, target, torch.zeros_like(target))
        w_t = (weight[safe] if weight is not None else torch.ones_like(pred[:, 0])) * keep
        picked = pred.gather(1, safe.unsqueeze(1)).squeeze(1)
        return -(picked * w_t).sum() / w_t.sum()
