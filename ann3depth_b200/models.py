"""Drop-in counterpart of the reference's `src/models.py`: module-level singletons `msdn` and `dcnf`
whose `__call__(images, depths, train=True)` builds the model once and returns ONE op
(src/models.py:179,277,370-371).  The driver's only interaction afterwards is to run that op
repeatedly (src/ann3depth.py:126-127) -- here `op.run()` (or `op()`).

images: float32 CUDA tensor [B,H,W,3] in [0,1]; depths: float32 CUDA tensor [B,h,w,1]; NHWC, static B
(src/data.py:82-86).  The tensors are the op's input buffers: refill them in place between runs.
Side channels the reference driver relies on are attributes of the op: `global_step`
(tf.train.get_or_create_global_step), `losses` (GraphKeys.LOSSES), `trainable_variables`.
"""
from __future__ import annotations

import torch

from . import ops
from .msdn import MSDNNet

_contexts = {}


def get_context(device=None) -> ops.Context:
    device = torch.cuda.current_device() if device is None else device
    if device not in _contexts:
        _contexts[device] = ops.Context(device)
    return _contexts[device]


def _check_inputs(images, depths, allow_u8=False):
    for t, ch, nm in ((images, 3, "images"), (depths, 1, "depths")):
        ok_dtype = t.dtype == torch.float32 or (allow_u8 and nm == "images" and t.dtype == torch.uint8)
        if not (torch.is_tensor(t) and t.is_cuda and ok_dtype and t.dim() == 4 and t.shape[-1] == ch
                and t.is_contiguous()):
            raise ValueError(f"{nm} must be a contiguous float32 CUDA tensor [B,H,W,{ch}] (NHWC)")
    if images.shape[0] != depths.shape[0]:
        raise ValueError("images and depths must have the same static batch size")


class MSDNTrainOp:
    """The object `models.msdn(images, depths)` returns: one `run()` == one `session.run(model_op)`."""

    def __init__(self, net: MSDNNet):
        self.net = net
        self.losses = {"loss/coarse_loss": net.loss_coarse, "loss/fine_loss": net.loss_fine}   # GraphKeys.LOSSES
        self.outputs = net.fine.view(net.B, 55, 74, 1)
        self.coarse = net.coarse.view(net.B, 55, 74, 1)

    @property
    def global_step(self):
        return self.net.global_step

    @property
    def trainable_variables(self):
        return {n: s.tf_shape for n, s in self.net.arena.specs.items()}

    def run(self, use_graph=True):
        if self.net.train:
            return self.net.train_step(use_graph=use_graph)
        return self.net.infer()

    __call__ = run


class _MultiScaleDeepNetwork:
    """Eigen et al. (2014) multi-scale deep network; mirrors src/models.py:203-367."""

    def __call__(self, images, depths, train=True, **net_kwargs):
        # beyond the reference contract: uint8 images (pixel values 0..255, read as pixel / 255) are accepted too --
        # the reference's own pixels are k/255 (tools/data_tf_converter.py:36-37), so this is the same data at a
        # quarter of the host->device bytes
        _check_inputs(images, depths, allow_u8=True)
        ctx = get_context(images.device.index)
        net = MSDNNet(ctx, images.shape[0], tuple(images.shape[1:3]), tuple(depths.shape[1:3]), train=train,
                      **net_kwargs)
        net.images, net.depths = images, depths        # the op reads its inputs from the caller's buffers
        return MSDNTrainOp(net)


class _DistributedConvolutionalNeuralFields:
    """Liu et al. (2015) deep convolutional neural field; mirrors src/models.py:9-200."""

    def __call__(self, images, depths, train=True, **net_kwargs):
        _check_inputs(images, depths)
        from .dcnf import DCNFNet, DCNFTrainOp
        ctx = get_context(images.device.index)
        net = DCNFNet(ctx, images.shape[0], tuple(images.shape[1:3]), tuple(depths.shape[1:3]), train=train,
                      **net_kwargs)
        net.images, net.depths = images, depths
        return DCNFTrainOp(net)


dcnf = _DistributedConvolutionalNeuralFields()   # src/models.py:370
msdn = _MultiScaleDeepNetwork()                  # src/models.py:371
