"""Drop-in ``SSD`` module: same constructor, attributes, ``state_dict`` keys and method names as the reference
``src/model/ssd.py``, with the detection-head math (default boxes, matching, encoding, MultiBox loss and its
gradient) executed by the hand-written sm_100a kernels of libssdhead.so.

The VGG16-bn trunk and the extra / detector convolutions stay on stock torch + cuDNN (out of scope, SURVEY
section 2); only their names are kept so reference checkpoints load unchanged.
"""
from __future__ import annotations

import pathlib
from typing import Iterator, Optional

import torch
import torch.nn as nn

from .. import ops

# (name of the activation a detector reads, in_channels, anchors per cell)       reference ssd.py:70-77
_DETECTOR_TAPS = (("4_3", 512, 4), ("7_1", 1024, 6), ("8_2", 512, 6), ("9_2", 256, 6), ("10_2", 256, 4), ("11_2", 256, 4))
# VGG16-bn trunk: convolution widths per stage; pooling after stages 1-4 (stage 3 pads, reference vgg16.py:26)
_VGG_STAGES = ((64, 64), (128, 128), (256, 256, 256), (512, 512, 512), (512, 512, 512))
# extra layers 6-11: (kernel, out_channels, stride, padding)                      reference ssd.py:49-54
_EXTRA_STAGES = (((3, 1024, 1, 1),), ((1, 1024, 1, 0),), ((1, 256, 1, 0), (3, 512, 2, 1)), ((1, 128, 1, 0), (3, 256, 2, 1)),
                 ((1, 128, 1, 0), (3, 256, 1, 0)), ((1, 128, 1, 0), (3, 256, 1, 0)))


class SSD(nn.Module):
    def __init__(self, num_classes: int, weights_path: Optional[pathlib.Path] = None,
                 weights_path_vgg16: Optional[pathlib.Path] = None):
        super().__init__()
        self.num_classes = num_classes
        # plain CPU tensor, not a buffer -- callers move it themselves (reference train.py:81, evaluate.py:110)
        self.default_bboxes = self._get_default_bboxes()

        feats = nn.ModuleDict()
        cin = 3
        for stage, widths in enumerate(_VGG_STAGES, start=1):
            for sub, width in enumerate(widths, start=1):
                self._add_block(feats, f"{stage}_{sub}", nn.Conv2d(cin, width, kernel_size=3, padding=1), frozen=True)
                cin = width
            if stage < 5:
                feats[f"pool_{stage}"] = nn.MaxPool2d(kernel_size=2, stride=2, padding=1 if stage == 3 else 0)
        for stage, convs in enumerate(_EXTRA_STAGES, start=6):
            for sub, (k, width, stride, pad) in enumerate(convs, start=1):
                self._add_block(feats, f"{stage}_{sub}", nn.Conv2d(cin, width, kernel_size=k, stride=stride, padding=pad), frozen=False)
                cin = width
        self.features = feats
        self.detectors = nn.ModuleDict({
            f"det_{tap}": nn.Conv2d(ch, anchors * (num_classes + 4), kernel_size=3, padding=1) for tap, ch, anchors in _DETECTOR_TAPS})

        if weights_path and pathlib.Path(weights_path).exists():
            self.load_state_dict(torch.load(pathlib.Path(weights_path).as_posix()))
        else:
            if weights_path_vgg16 and pathlib.Path(weights_path_vgg16).exists():
                self._load_vgg16_trunk(torch.load(pathlib.Path(weights_path_vgg16).as_posix()))
            self._initialize_weights()

    @staticmethod
    def _add_block(feats: nn.ModuleDict, tag: str, conv: nn.Conv2d, frozen: bool) -> None:
        bn = nn.BatchNorm2d(conv.out_channels)
        if frozen:                                   # reference ssd.py:31-32 freezes the VGG trunk
            for prm in list(conv.parameters()) + list(bn.parameters()):
                prm.requires_grad = False
        feats[f"conv_{tag}"] = conv
        feats[f"bn_{tag}"] = bn
        feats[f"act_{tag}"] = nn.ReLU(inplace=True)

    def _load_vgg16_trunk(self, vgg_state: dict) -> None:
        """Map a reference VGG16 checkpoint (``features.<n>.*``, vgg16.py:24-38) onto the named trunk layers."""
        names = [k for k in self.features.keys() if int(k.split("_")[1]) <= 5]
        # the reference Sequential also holds the fifth pooling layer, which SSD drops (ssd.py:37-41)
        seq_index = {}
        n = 0
        for stage, widths in enumerate(_VGG_STAGES, start=1):
            for sub in range(1, len(widths) + 1):
                for kind in ("conv", "bn", "act"):
                    seq_index[f"{kind}_{stage}_{sub}"] = n
                    n += 1
            n += 1                                   # the pooling slot
        own = self.state_dict()
        for name in names:
            if name.startswith(("act", "pool")):
                continue
            for suffix in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked"):
                src = f"features.{seq_index[name]}.{suffix}"
                dst = f"features.{name}.{suffix}"
                if src in vgg_state and dst in own:
                    own[dst].copy_(vgg_state[src])

    def normalize(self, x: torch.Tensor) -> torch.Tensor:
        """ImageNet mean / std normalisation (reference vgg16.py:103-115)."""
        mean = torch.tensor([0.485, 0.456, 0.406], device=x.device, dtype=x.dtype).view(1, 3, 1, 1)
        std = torch.tensor([0.229, 0.224, 0.225], device=x.device, dtype=x.dtype).view(1, 3, 1, 1)
        return x.sub(mean).div(std)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """(N, 3, 300, 300) -> (N, 8732, 4 + C), rows level-major then (cell row, cell col, anchor).

        The reference looks detectors up with the activation's name and therefore finds none (ssd.py:102, SURVEY
        section 0 item 3); the evidently intended ``act_*`` -> ``det_*`` pairing is used here.  The six detector outputs go through
        ``ssdh_pack_head``: permute + reshape + cat (ssd.py:103-104) as one pass over the data."""
        n = x.size(0)
        width = self.num_classes + 4
        x = self.normalize(x)
        levels = []
        for name, layer in self.features.items():
            x = layer(x)
            det = "det" + name[3:] if name.startswith("act") else None
            if det is not None and det in self.detectors:
                levels.append(self.detectors[det](x))
        if not levels:
            return x.new_empty((n, 0, width))
        # no host path: the tail is ssdh_pack_head (permute + reshape + cat of ssd.py:103-104 as one pass) and raises on CPU
        # tensors like every other entry point of the head
        return ops.pack_head(levels, width)

    def _get_default_bboxes(self) -> torch.Tensor:
        """(8732, 4) priors; kernel ssdh_default_boxes, bit-identical to reference ssd.py:108-133."""
        return ops.default_boxes("cuda").cpu()

    def _initialize_weights(self) -> None:
        trainable = [m for k, m in self.features.items() if int(k.split("_")[1]) >= 6] + list(self.detectors.values())
        for m in trainable:
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def train_params(self) -> Iterator[nn.Parameter]:
        """Extra layers (6-11) then detectors, the order Adam sees them in reference train.py:97."""
        for name, layer in self.features.items():
            if int(name.split("_")[1]) >= 6:
                yield from layer.parameters()
        for layer in self.detectors.values():
            yield from layer.parameters()

    # ---- head math: thin wrappers over the kernels ------------------------------------------------------------
    def loss(self, outputs: torch.Tensor, targets: torch.Tensor, default_bboxes: torch.Tensor, a: int = 1,
             force_best_prior: bool = False) -> torch.Tensor:
        """MultiBox loss, 0-dim, differentiable w.r.t. ``outputs`` (reference ssd.py:181-229): one fused launch.
        ``force_best_prior`` is north_star's opt-in extension of the matching; off = the reference."""
        return ops.multibox_loss(outputs, targets, default_bboxes, a=a, force_best_prior=force_best_prior)

    def _match(self, gt: torch.Tensor, df: torch.Tensor, threshold: float = 0.25, force_best_prior: bool = False) -> torch.Tensor:
        return ops.match(gt, df, threshold, force_best_prior=force_best_prior).mask

    def _calc_delta(self, gt: torch.Tensor, df: torch.Tensor) -> torch.Tensor:
        return ops.encode(gt, df)

    def _smooth_l1(self, x: torch.Tensor) -> torch.Tensor:
        return ops.smooth_l1(x)

    def _softmax_cross_entropy(self, pr: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
        return ops.softmax_cross_entropy(pr, gt)

    def _split_pos_neg(self, pos_num: torch.Tensor, neg_num: torch.Tensor) -> tuple:
        return ops.split_pos_neg(pos_num, neg_num)

    def _k_plus_1_th_value(self, tensor: torch.Tensor, k: torch.Tensor) -> torch.Tensor:
        return ops.kplus1_value(tensor, k)
