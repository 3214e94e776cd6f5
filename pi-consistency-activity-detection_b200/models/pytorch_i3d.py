"""Drop-in for the reference's models/pytorch_i3d.py (MaxPool3dSamePadding :13-45, Unit3D :48-120,
InceptionModule :124-149, InceptionI3d :152-346).  Same classes / ctor args / state_dict keys; the
forward passes run on the b200caps sm_100a kernels (tcgen05 implicit-GEMM conv, fused BN/ReLU/pool).

Tensors crossing module boundaries are logical (N,C,T,H,W); internally they are bf16 channels-last views
(zero-copy between modules).  fp32 NCDHW inputs are accepted and converted by a layout kernel.
CPU tensors are rejected: there is no fallback path.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from b200caps import engine
from b200caps.plans import ConvSpec, same_pad


class MaxPool3dSamePadding(nn.MaxPool3d):
    """TF-style 'same' zero padding then max-pool (reference :13-45)."""

    def compute_pad(self, dim, s):
        if s % self.stride[dim] == 0:
            return max(self.kernel_size[dim] - self.stride[dim], 0)
        return max(self.kernel_size[dim] - (s % self.stride[dim]), 0)

    def forward(self, x):
        x_cl = engine.to_cl(x)
        y = engine.MaxPoolFn.apply(x_cl, tuple(self.kernel_size), tuple(self.stride))
        return engine.from_cl(y)


class Unit3D(nn.Module):
    """same-pad Conv3d (no bias) -> BatchNorm3d(eps=1e-3, momentum=0.01) -> ReLU (reference :48-120)."""

    def __init__(self, in_channels, output_channels, kernel_shape=(1, 1, 1), stride=(1, 1, 1), padding=0,
                 activation_fn=F.relu, use_batch_norm=True, use_bias=False, name='unit_3d'):
        super(Unit3D, self).__init__()
        self._output_channels = output_channels
        self._kernel_shape = tuple(kernel_shape)
        self._stride = tuple(stride)
        self._use_batch_norm = use_batch_norm
        self._activation_fn = activation_fn
        self._use_bias = use_bias
        self.name = name
        self.padding = padding
        self._in_channels = in_channels
        # parameter holders with the reference's names; their own forward is never called
        self.conv3d = nn.Conv3d(in_channels=in_channels, out_channels=output_channels, kernel_size=self._kernel_shape,
                                stride=self._stride, padding=0, bias=self._use_bias)
        if self._use_batch_norm:
            self.bn = nn.BatchNorm3d(self._output_channels, eps=0.001, momentum=0.01)
        self.__dict__["_layer_cache"] = None

    def compute_pad(self, dim, s):
        if s % self._stride[dim] == 0:
            return max(self._kernel_shape[dim] - self._stride[dim], 0)
        return max(self._kernel_shape[dim] - (s % self._stride[dim]), 0)

    @property
    def _layer(self):
        lc = self.__dict__["_layer_cache"]
        if lc is None or lc.weight is not self.conv3d.weight:
            k, s = self._kernel_shape, self._stride
            cin, cout = self._in_channels, self._output_channels

            def spec_fn(dims):
                pads = [same_pad(d, kk, ss) for d, kk, ss in zip(dims, k, s)]
                return ConvSpec(cin, cout, k, s, tuple(p[0] for p in pads), tuple(p[1] for p in pads))

            if cin < 8 and k[0] * k[1] * k[2] > 1:
                lc = engine.StemLayer(self.conv3d.weight, cin, cout, k, s)   # RGB stem: explicit im2col + TMA GEMM
            else:
                lc = engine.ConvLayer(self.conv3d.weight, spec_fn)
            self.__dict__["_layer_cache"] = lc
        return lc

    def forward(self, x):
        if not (self._use_batch_norm and self._activation_fn is F.relu and not self._use_bias):
            raise NotImplementedError("b200caps Unit3D implements the conv+BN+ReLU form used up to Mixed_4f "
                                      "(the truncated trunk of capsules_ucf101.py:343); the Logits head is outside the hot path")
        x_cl = engine.to_cl(x, (self._in_channels + 7) // 8 * 8)
        y = engine.Unit3DFn.apply(x_cl, self.conv3d.weight, self.bn.weight, self.bn.bias, self)
        return engine.from_cl(y)


class InceptionModule(nn.Module):
    """Four branches + channel concat (reference :124-149), executed as one hand-scheduled function."""

    def __init__(self, in_channels, out_channels, name):
        super(InceptionModule, self).__init__()
        self.b0 = Unit3D(in_channels=in_channels, output_channels=out_channels[0], kernel_shape=[1, 1, 1], padding=0,
                         name=name + '/Branch_0/Conv3d_0a_1x1')
        self.b1a = Unit3D(in_channels=in_channels, output_channels=out_channels[1], kernel_shape=[1, 1, 1], padding=0,
                          name=name + '/Branch_1/Conv3d_0a_1x1')
        self.b1b = Unit3D(in_channels=out_channels[1], output_channels=out_channels[2], kernel_shape=[3, 3, 3],
                          name=name + '/Branch_1/Conv3d_0b_3x3')
        self.b2a = Unit3D(in_channels=in_channels, output_channels=out_channels[3], kernel_shape=[1, 1, 1], padding=0,
                          name=name + '/Branch_2/Conv3d_0a_1x1')
        self.b2b = Unit3D(in_channels=out_channels[3], output_channels=out_channels[4], kernel_shape=[3, 3, 3],
                          name=name + '/Branch_2/Conv3d_0b_3x3')
        self.b3a = MaxPool3dSamePadding(kernel_size=[3, 3, 3], stride=(1, 1, 1), padding=0)
        self.b3b = Unit3D(in_channels=in_channels, output_channels=out_channels[5], kernel_shape=[1, 1, 1], padding=0,
                          name=name + '/Branch_3/Conv3d_0b_1x1')
        self.name = name

    def forward(self, x):
        x_cl = engine.to_cl(x)
        params = []
        for n in engine.InceptionFn.UNITS:
            u = getattr(self, n)
            params += [u.conv3d.weight, u.bn.weight, u.bn.bias]
        y = engine.InceptionFn.apply(x_cl, self, *params)
        return engine.from_cl(y)


class InceptionI3d(nn.Module):
    """Inception-v1 I3D (reference :152-346).  forward returns (x, out56, out112) like the reference."""

    VALID_ENDPOINTS = (
        'Conv3d_1a_7x7', 'MaxPool3d_2a_3x3', 'Conv3d_2b_1x1', 'Conv3d_2c_3x3', 'MaxPool3d_3a_3x3', 'Mixed_3b', 'Mixed_3c',
        'MaxPool3d_4a_3x3', 'Mixed_4b', 'Mixed_4c', 'Mixed_4d', 'Mixed_4e', 'Mixed_4f', 'MaxPool3d_5a_2x2', 'Mixed_5b',
        'Mixed_5c', 'Logits', 'Predictions',
    )

    def __init__(self, num_classes=400, spatial_squeeze=True, final_endpoint='Logits', name='inception_i3d',
                 in_channels=3, dropout_keep_prob=0.5):
        if final_endpoint not in self.VALID_ENDPOINTS:
            raise ValueError('Unknown final endpoint %s' % final_endpoint)
        super(InceptionI3d, self).__init__()
        self._num_classes = num_classes
        self._spatial_squeeze = spatial_squeeze
        self._final_endpoint = final_endpoint
        self.logits = None
        self.end_points = {}
        table = [
            ('Conv3d_1a_7x7', lambda: Unit3D(in_channels=in_channels, output_channels=64, kernel_shape=[7, 7, 7],
                                             stride=(2, 2, 2), padding=(3, 3, 3), name=name + 'Conv3d_1a_7x7')),
            ('MaxPool3d_2a_3x3', lambda: MaxPool3dSamePadding(kernel_size=[1, 3, 3], stride=(1, 2, 2), padding=0)),
            ('Conv3d_2b_1x1', lambda: Unit3D(in_channels=64, output_channels=64, kernel_shape=[1, 1, 1], padding=0,
                                             name=name + 'Conv3d_2b_1x1')),
            ('Conv3d_2c_3x3', lambda: Unit3D(in_channels=64, output_channels=192, kernel_shape=[3, 3, 3],
                                             stride=(2, 1, 1), padding=1, name=name + 'Conv3d_2c_3x3')),
            ('MaxPool3d_3a_3x3', lambda: MaxPool3dSamePadding(kernel_size=[1, 3, 3], stride=(1, 2, 2), padding=0)),
            ('Mixed_3b', lambda: InceptionModule(192, [64, 96, 128, 16, 32, 32], name + 'Mixed_3b')),
            ('Mixed_3c', lambda: InceptionModule(256, [128, 128, 192, 32, 96, 64], name + 'Mixed_3c')),
            ('MaxPool3d_4a_3x3', lambda: MaxPool3dSamePadding(kernel_size=[3, 3, 3], stride=(2, 1, 1), padding=0)),
            ('Mixed_4b', lambda: InceptionModule(128 + 192 + 96 + 64, [192, 96, 208, 16, 48, 64], name + 'Mixed_4b')),
            ('Mixed_4c', lambda: InceptionModule(192 + 208 + 48 + 64, [160, 112, 224, 24, 64, 64], name + 'Mixed_4c')),
            ('Mixed_4d', lambda: InceptionModule(160 + 224 + 64 + 64, [128, 128, 256, 24, 64, 64], name + 'Mixed_4d')),
            ('Mixed_4e', lambda: InceptionModule(128 + 256 + 64 + 64, [112, 144, 288, 32, 64, 64], name + 'Mixed_4e')),
            ('Mixed_4f', lambda: InceptionModule(112 + 288 + 64 + 64, [256, 160, 320, 32, 128, 128], name + 'Mixed_4f')),
            ('MaxPool3d_5a_2x2', lambda: MaxPool3dSamePadding(kernel_size=[2, 2, 2], stride=(2, 2, 2), padding=0)),
            ('Mixed_5b', lambda: InceptionModule(256 + 320 + 128 + 128, [256, 160, 320, 32, 128, 128], name + 'Mixed_5b')),
            ('Mixed_5c', lambda: InceptionModule(256 + 320 + 128 + 128, [384, 192, 384, 48, 128, 128], name + 'Mixed_5c')),
        ]
        done = False
        for end_point, make in table:
            self.end_points[end_point] = make()
            if self._final_endpoint == end_point:
                done = True
                break
        if done:
            self.build()
            return
        # full network to Logits (reference :297-310); outside the hot path, kept for constructor parity
        self.avg_pool = nn.AvgPool3d(kernel_size=[2, 7, 7], stride=(1, 1, 1))
        self.dropout = nn.Dropout(dropout_keep_prob)
        self.logits = Unit3D(in_channels=384 + 384 + 128 + 128, output_channels=self._num_classes, kernel_shape=[1, 1, 1],
                             padding=0, activation_fn=None, use_batch_norm=False, use_bias=True, name='logits')
        self.build()

    def replace_logits(self, num_classes):
        self._num_classes = num_classes
        self.logits = Unit3D(in_channels=384 + 384 + 128 + 128, output_channels=self._num_classes, kernel_shape=[1, 1, 1],
                             padding=0, activation_fn=None, use_batch_norm=False, use_bias=True, name='logits')

    def build(self):
        for k in list(self.end_points.keys()):
            self.add_module(k, self.end_points[k])

    def forward(self, x):
        out56 = out112 = None
        for end_point in self.VALID_ENDPOINTS:
            if end_point in self.end_points:
                x = self._modules[end_point](x)
                if end_point == 'Conv3d_2c_3x3':
                    out56 = x
                if end_point == 'Conv3d_1a_7x7':
                    out112 = x
        return (x, out56, out112)

    def extract_features(self, x):
        for end_point in self.VALID_ENDPOINTS:
            if end_point in self.end_points:
                x = self._modules[end_point](x)
        return self.avg_pool(x)
