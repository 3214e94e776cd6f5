"""Drop-in for the reference's models/capsules_ucf101.py (PrimaryCaps :10-49, ConvCaps :52-331,
CapsNet :334-512): same class names, constructor arguments, forward signatures, return shapes and
state_dict keys.  Everything numeric runs in the b200caps sm_100a kernels:

  * PrimaryCaps          one tcgen05 implicit GEMM (pose|a fused, N = 544, sigmoid + NHWC in the epilogue)
  * ConvCaps             fused EM routing kernel (3 iterations, votes kept in registers), hand-derived backward
  * decoder              transposed convs by output-parity class, skips written into concat slots,
                         Dropout3d folded into the upsample4 epilogue, `smooth` = projection GEMM + stencil
"""
import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from b200caps import engine
from b200caps.plans import ConvSpec
from models.pytorch_i3d import InceptionI3d


class PrimaryCaps(nn.Module):
    r"""Primary convolutional capsule layer (reference :10-49).
    input (*, A, h, w) -> output (*, h', w', B*(P*P+1)) (poses first, then activations)."""

    def __init__(self, A, B, K, P, stride):
        super(PrimaryCaps, self).__init__()
        self.pose = nn.Conv2d(in_channels=A, out_channels=B * P * P, kernel_size=K, stride=stride, bias=True)
        self.pose.weight.data.normal_(0.0, 0.1)
        self.a = nn.Conv2d(in_channels=A, out_channels=B, kernel_size=K, stride=stride, bias=True)
        self.a.weight.data.normal_(0.0, 0.1)
        self.sigmoid = nn.Sigmoid()
        self._A, self._B, self._K, self._PP, self._stride = A, B, K, P * P, stride
        self.__dict__["_layer_cache"] = None

    @property
    def _layer(self):
        lc = self.__dict__["_layer_cache"]
        if lc is None or lc.weights[0] is not self.pose.weight:
            A, K, n_out, st = self._A, self._K, self._B * (self._PP + 1), self._stride

            def spec_fn(dims):
                return ConvSpec(A, n_out, (1, K, K), (1, st, st))

            lc = engine.FusedConvLayer([self.pose.weight, self.a.weight], spec_fn, grad_cpad=576, fprop_ksplit=engine.PC_KSPLIT)
            self.__dict__["_layer_cache"] = lc
        return lc

    def forward(self, x):
        if self._B * self._PP != 512 or self._B != 32:
            raise NotImplementedError("b200caps PrimaryCaps is specialised for B=32, P=4 (capsules_ucf101.py:355)")
        x_cl = engine.to_cl(x)
        out = engine.PrimaryCapsFn.apply(x_cl, self.pose.weight, self.pose.bias, self.a.weight, self.a.bias, self)
        return out[:, 0]           # (N, h', w', 544) fp32


class ConvCaps(nn.Module):
    r"""Convolutional capsule layer with EM routing (reference :52-331).
    input (*, h, w, B*(P*P+1)) -> output (*, h', w', C*(P*P+1)).  Only the configuration the model uses is
    implemented (K=(1,1), stride (1,1), w_shared=False, coor_add=False, B=32, P=4; reference :356)."""

    def __init__(self, B, C, K, P, stride, iters=3, coor_add=False, w_shared=False):
        super(ConvCaps, self).__init__()
        self.B, self.C, self.K, self.P = B, C, K, P
        self.psize = P * P
        self.stride = stride
        self.iters = iters
        self.coor_add = coor_add
        self.w_shared = w_shared
        self.eps = 1e-8
        self._lambda = 1e-6
        self.beta_u = nn.Parameter(torch.randn(C, self.psize))
        self.beta_a = nn.Parameter(torch.randn(C))
        self.weights = nn.Parameter(torch.randn(1, K[0] * K[1] * B, C, P, P))
        self.sigmoid = nn.Sigmoid()
        self.softmax = nn.Softmax(dim=2)

    def forward(self, x):
        if (tuple(self.K) != (1, 1) or tuple(self.stride) != (1, 1) or self.w_shared or self.coor_add or self.B != 32
                or self.P != 4 or self.iters != 3 or self.C > 32):
            raise NotImplementedError("b200caps ConvCaps implements the routing configuration of capsules_ucf101.py:356")
        engine.require_cuda(x, "capsules")
        return engine.EMRoutingFn.apply(x.float(), self.weights, self.beta_u, self.beta_a)


class CapsNet(nn.Module):
    NUM_CLASSES = 24

    def __init__(self, pt_path='../weights/rgb_charades.pt', P=4, pretrained_load='i3d'):
        super(CapsNet, self).__init__()
        self.P = P
        C = self.NUM_CLASSES
        self.conv1 = InceptionI3d(157, in_channels=3, final_endpoint='Mixed_4f')
        if pt_path is not None and os.path.isfile(pt_path):
            pretrained_weights = torch.load(pt_path, map_location="cpu")
            weights = self.conv1.state_dict()
            loaded_layers = 0
            for name in weights.keys():
                if name in pretrained_weights.keys():
                    weights[name] = pretrained_weights[name]
                    loaded_layers += 1
            self.conv1.load_state_dict(weights)
            print("Loaded I3D pretrained weights from ", pt_path, " for layers: ", loaded_layers)
        else:
            print("I3D pretrained weights not found at ", pt_path, " -- encoder keeps its random initialisation")

        self.primary_caps = PrimaryCaps(832, 32, 9, P, stride=1)
        self.conv_caps = ConvCaps(32, C, (1, 1), P, stride=(1, 1), iters=3)

        self.upsample1 = nn.ConvTranspose2d(C * P * P, 64, kernel_size=9, stride=1, padding=0)
        self.upsample1.weight.data.normal_(0.0, 0.02)
        self.upsample2 = nn.ConvTranspose3d(128, 64, kernel_size=(3, 3, 3), stride=(2, 2, 2), padding=1, output_padding=1)
        self.upsample2.weight.data.normal_(0.0, 0.02)
        self.upsample3 = nn.ConvTranspose3d(128, 64, kernel_size=(3, 3, 3), stride=(2, 2, 2), padding=1, output_padding=1)
        self.upsample3.weight.data.normal_(0.0, 0.02)
        self.upsample4 = nn.ConvTranspose3d(128, 128, kernel_size=(3, 3, 3), stride=(2, 2, 2), padding=1,
                                            output_padding=(1, 1, 1))
        self.upsample4.weight.data.normal_(0.0, 0.02)
        self.dropout3d = nn.Dropout3d(0.5)
        self.smooth = nn.ConvTranspose3d(128, 1, kernel_size=3, padding=1)
        self.smooth.weight.data.normal_(0.0, 0.02)
        self.relu = nn.ReLU()
        self.sig = nn.Sigmoid()
        self.conv28 = nn.Conv2d(832, 64, kernel_size=(3, 3), padding=(1, 1))
        self.conv56 = nn.Conv3d(192, 64, kernel_size=(3, 3, 3), padding=(1, 1, 1))
        self.conv112 = nn.Conv3d(64, 64, kernel_size=(3, 3, 3), padding=(1, 1, 1))
        self.__dict__["_layers_cache"] = None

    # ---- derived kernel state -------------------------------------------------------------------
    @property
    def _layers(self):
        lc = self.__dict__["_layers_cache"]
        if lc is None or lc["upsample1"].weight is not self.upsample1.weight:
            C16 = self.NUM_CLASSES * self.P * self.P
            one, two, z = (1, 1, 1), (2, 2, 2), (0, 0, 0)
            lc = {
                "upsample1": engine.ConvLayer(self.upsample1.weight,
                                              lambda d: ConvSpec(C16, 64, (1, 9, 9), one, z, z, z, True)),
                "conv28": engine.ConvLayer(self.conv28.weight, lambda d: ConvSpec(832, 64, (1, 3, 3), one, (0, 1, 1), (0, 1, 1))),
                "upsample2": engine.ConvLayer(self.upsample2.weight,
                                              lambda d: ConvSpec(128, 64, (3, 3, 3), two, one, z, one, True)),
                "conv56": engine.ConvLayer(self.conv56.weight, lambda d: ConvSpec(192, 64, (3, 3, 3), one, one, one)),
                "upsample3": engine.ConvLayer(self.upsample3.weight,
                                              lambda d: ConvSpec(128, 64, (3, 3, 3), two, one, z, one, True)),
                "conv112": engine.ConvLayer(self.conv112.weight, lambda d: ConvSpec(64, 64, (3, 3, 3), one, one, one)),
                "upsample4": engine.ConvLayer(self.upsample4.weight,
                                              lambda d: ConvSpec(128, 128, (3, 3, 3), two, one, z, one, True)),
                "smooth": engine.SmoothLayer(self.smooth.weight),
                "tail": engine.CollapsedTail(self.upsample4, self.smooth),
            }
            self.__dict__["_layers_cache"] = lc
        return lc

    def load_pretrained_weights(self):
        saved_weights = torch.load('./savedweights/weights_referit')
        self.load_state_dict(saved_weights, strict=False)
        print('loaded referit pretrained weights for whole network')

    def load_previous_weights(self, weightfile):
        saved_weights = torch.load(weightfile)
        self.load_state_dict(saved_weights, strict=False)
        print('loaded weights from previous run: ', weightfile)

    def caps_reorder(self, imgcaps):
        """Identity re-concatenation of poses and activations (reference :398-410)."""
        num_imgcaps = int(imgcaps.size()[3] / (self.P * self.P))
        pose_range = num_imgcaps * self.P * self.P
        return torch.cat((imgcaps[:, :, :, :pose_range], imgcaps[:, :, :, pose_range:pose_range + num_imgcaps]), dim=-1)

    # The forward pass in three segments (tests drive them separately: the EM routing in the middle is
    # chaotically sensitive at random init, so parity is established per segment).
    def _encode(self, img):
        """I3D trunk + Dropout3d -> (x (N,1,28,28,832) bf16 CL, cross56 CL, cross112 CL, decoder dropout scale)."""
        x, cross56, cross112 = self.conv1(img)
        x_cl = engine.to_cl(x)
        N, dev = x_cl.shape[0], x_cl.device
        drop2 = None
        if self.training:
            # nn.Dropout3d(0.5): one Bernoulli per (sample, channel) (reference :428, :507)
            x_cl = engine.ChannelScaleFn.apply(x_cl, engine.dropout_scale(N, 832, dev))
            drop2 = engine.dropout_scale(N, 128, dev)
        return x_cl, engine.to_cl(cross56), engine.to_cl(cross112), drop2

    def _capsules(self, x_cl):
        """PrimaryCaps GEMM -> (N,20,20,544) fp32 ; EM routing -> (N,20,20,C*17) fp32 [mu | a]."""
        pc = self.primary_caps
        caps = engine.PrimaryCapsFn.apply(x_cl, pc.pose.weight, pc.pose.bias, pc.a.weight, pc.a.bias, pc)[:, 0]
        rout = engine.EMRoutingFn.apply(caps, self.conv_caps.weights, self.conv_caps.beta_u, self.conv_caps.beta_a)
        return caps, rout

    def _decode(self, rout, x_cl, c56, c112, drop2, classification, concat_labels, epoch, thresh_ep):
        C = self.NUM_CLASSES
        N, h, w = rout.shape[0], rout.shape[1], rout.shape[2]
        dev = rout.device
        actor_prediction = engine.ClassActFn.apply(rout)
        feat_shape = rout[..., C * 16:].reshape(N, h * w, C)
        # pose mask (reference :455-479), built on the device without host round trips
        with torch.no_grad():
            if self.training:
                lab = F.one_hot(classification.to(dev).long().view(-1), C).float()
                if epoch < thresh_ep:
                    unl = torch.ones_like(lab)
                else:
                    unl = F.one_hot(torch.argmax(actor_prediction, dim=1), C).float()
                sel = (concat_labels.to(dev).view(-1, 1) == 0).float()
                mask = (sel * unl + (1.0 - sel) * lab).contiguous()
            else:
                mask = F.one_hot(torch.argmax(actor_prediction, dim=1), C).float().contiguous()
        x0 = engine.CapsHeadFn.apply(rout, mask)
        params = []
        for n in engine.DecoderFn.ORDER:
            m = getattr(self, n)
            params += [m.weight, m.bias]
        out_1 = engine.DecoderFn.apply(x0, x_cl, c56, c112, drop2, self, *params)
        return out_1, actor_prediction, feat_shape

    def forward(self, img, classification, concat_labels, epoch, thresh_ep):
        '''
        img (B,3,T,H,W); classification (B,1) ground-truth class (pose masking of labeled clips);
        concat_labels (B,) 1 = labeled, 0 = unlabeled.
        Returns (mask logits (B,1,8,224,224) fp32, class activations (B,C) fp32, per-location activations (B,400,C) fp32).
        '''
        engine.require_cuda(img, "img")
        x_cl, c56, c112, drop2 = self._encode(img)
        caps, rout = self._capsules(x_cl)
        return self._decode(rout, x_cl, c56, c112, drop2, classification, concat_labels, epoch, thresh_ep)
