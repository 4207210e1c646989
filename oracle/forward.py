"""CPU restatement of the VNect graph (reference: src/vnect_model.py:27-217) on torch-CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).  TensorFlow 1.x defaults that the reference relies on and that are
restated here ([TF1 API]):

* ``tc.layers.conv2d``: conv + bias + ReLU unless ``activation_fn=None``; ``padding='same'`` is TF SAME
  (``out = ceil(in/s)``, ``pad_total = max((out-1)*s + k - in, 0)``, ``before = pad_total // 2``).
* ``tc.layers.max_pool2d(kernel_size=3)``: stride 2 (vnect_model.py:29), SAME, padded cells ignored.
* ``tc.layers.batch_norm(scale=True, is_training=False)``: ``gamma * (x - mean) / sqrt(var + 0.001) + beta``.
* ``tf.layers.conv2d_transpose(kernel_size=4, strides=2, padding='same', use_bias=False)``: gradient-of-conv,
  kernel [kh, kw, out, in], ``X[i] = sum_{o,k: 2o+k-1=i} Y[o] W[k]`` == ``torch.conv_transpose2d(stride=2, padding=1)``.
* The wiring at vnect_model.py:56 feeds ``res2b_branch2a`` (not ``res2c_branch2a``) into ``res2c_branch2b``; it is
  reproduced, so ``res2c_branch2a`` is dead.

Parity note: the CNN restatement is "parity unpinned" against TensorFlow (not installable, no goldens in the
reference); its shapes are pinned against materials/caffe_script.ipynb in tests/test_oracle_forward.py.
"""
import numpy as np
import torch
import torch.nn.functional as F


def _same_pad(size, k, s):
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


class OracleNet:
    """Callable: NHWC float32 [n,S,S,3] -> (heatmap, x, y, z) each NHWC [n,S/8,S/8,21], plus optional layer taps."""

    def __init__(self, weights, dtype=torch.float32, threads=None):
        self.dtype = dtype
        if threads:
            torch.set_num_threads(threads)
        self.w = {}
        for name, arr in weights.items():
            t = torch.from_numpy(np.ascontiguousarray(arr)).to(dtype)
            leaf = name.split("/")[1]
            if leaf == "weights" or name == "res5c_branch2c/kernel":
                t = t.permute(3, 2, 0, 1).contiguous()  # HWIO -> OIHW
            elif leaf == "kernel":
                t = t.permute(3, 2, 0, 1).contiguous()  # [kh,kw,out,in] -> [in,out,kh,kw] (conv_transpose2d layout)
            self.w[name] = t

    # -- primitive layers --------------------------------------------------------------------------------------
    def _conv(self, x, scope, stride=1, relu=True, padding="valid"):
        w = self.w[scope + "/weights"]
        b = self.w[scope + "/biases"]
        k = w.shape[2]
        if padding == "same" and k > 1:
            pt, pb = _same_pad(x.shape[2], k, stride)
            pl, pr = _same_pad(x.shape[3], k, stride)
            x = F.pad(x, (pl, pr, pt, pb))
        y = F.conv2d(x, w, b, stride=stride)
        return F.relu(y) if relu else y

    def _block(self, x, prefix, proj, stride=1, suffix="", taps=None, a_input=None):
        """Bottleneck: 1x1 (stride) -> 3x3 -> 1x1, plus identity or 1x1 projection shortcut, add, ReLU."""
        short = self._conv(x, f"{prefix}_branch1{suffix}", stride=stride, relu=False) if proj else x
        a = self._conv(x, f"{prefix}_branch2a{suffix}", stride=stride) if a_input is None else a_input
        b = self._conv(a, f"{prefix}_branch2b{suffix}", padding="same")
        c = self._conv(b, f"{prefix}_branch2c{suffix}", relu=False)
        out = F.relu(c + short)
        if taps is not None:
            taps[f"{prefix}_branch2a{suffix}"] = a
            taps[f"{prefix}_branch2b{suffix}"] = b
            taps[prefix] = out
        return out, a

    @torch.no_grad()
    def forward(self, nhwc, want_taps=False):
        taps = {} if want_taps else None
        x = torch.from_numpy(np.ascontiguousarray(nhwc)).to(self.dtype).permute(0, 3, 1, 2).contiguous()
        x = self._conv(x, "conv1", stride=2, padding="same")  # vnect_model.py:27
        if taps is not None:
            taps["conv1"] = x
        pt, pb = _same_pad(x.shape[2], 3, 2)
        pl, pr = _same_pad(x.shape[3], 3, 2)
        x = F.max_pool2d(F.pad(x, (pl, pr, pt, pb), value=float("-inf")), 3, 2)  # :29
        if taps is not None:
            taps["pool1"] = x
        x, _ = self._block(x, "res2a", True, taps=taps)  # :32-41
        x, a2b = self._block(x, "res2b", False, taps=taps)  # :44-51
        x, _ = self._block(x, "res2c", False, taps=taps, a_input=a2b)  # :54-61 (branch2b reads res2b_branch2a)
        x, _ = self._block(x, "res3a", True, stride=2, taps=taps)  # :64-73
        for b in "bcd":
            x, _ = self._block(x, "res3" + b, False, taps=taps)  # :76-103
        x, _ = self._block(x, "res4a", True, stride=2, taps=taps)  # :106-115
        for b in "bcdef":
            x, _ = self._block(x, "res4" + b, False, taps=taps)  # :118-165
        x, _ = self._block(x, "res5a", True, suffix="_new", taps=taps)  # :168-177
        x = self._conv(x, "res5b_branch2a_new")  # :180
        x = self._conv(x, "res5b_branch2b_new", padding="same")  # :182
        x = self._conv(x, "res5b_branch2c_new")  # :184 (default activation: ReLU)
        if taps is not None:
            taps["res5b_branch2c_new"] = x
        d1 = F.conv_transpose2d(x, self.w["res5c_branch1a/kernel"], stride=2, padding=1)  # :188
        d2 = F.conv_transpose2d(x, self.w["res5c_branch2a/kernel"], stride=2, padding=1)  # :191
        g, be = self.w["bn5c_branch2a/gamma"], self.w["bn5c_branch2a/beta"]
        mu, var = self.w["bn5c_branch2a/moving_mean"], self.w["bn5c_branch2a/moving_variance"]
        d2 = (d2 - mu[None, :, None, None]) * (g / torch.sqrt(var + 0.001))[None, :, None, None] + be[None, :, None, None]
        d2 = F.relu(d2)  # :194-196
        sq = d1 * d1
        bone = torch.sqrt(sq[:, 0:21] + sq[:, 21:42] + sq[:, 42:63])  # :198-205
        feat = torch.cat([d2, d1[:, 0:21], d1[:, 21:42], d1[:, 42:63], bone], dim=1)  # :207-209
        if taps is not None:
            taps["res5c_branch2a_feat"] = feat
        y = self._conv(feat, "res5c_branch2b", padding="same")  # :211
        if taps is not None:
            taps["res5c_branch2b"] = y
        y = F.conv2d(y, self.w["res5c_branch2c/kernel"])  # :213 (no bias, linear)
        y = y.permute(0, 2, 3, 1).contiguous().to(torch.float32).numpy()
        outs = tuple(np.ascontiguousarray(y[..., 21 * i:21 * (i + 1)]) for i in range(4))  # :216
        if want_taps:
            return outs, {k: v.permute(0, 2, 3, 1).contiguous().to(torch.float32).numpy() for k, v in taps.items()}
        return outs

    __call__ = forward


def algorithmic_flops(box_size=368):
    """2 x live MACs of one forward: dead res2c_branch2a excluded, deconv = 4 taps x Cin."""
    from .weights import CONV_SCOPES
    s = box_size
    hw = {"conv1": (s // 2) ** 2}
    macs = 0
    for scope, k, cin, cout in CONV_SCOPES:
        if scope == "res2c_branch2a":
            continue
        if scope == "conv1":
            px = (s // 2) ** 2
        elif scope.startswith("res2"):
            px = (s // 4) ** 2
        elif scope.startswith("res3"):
            px = (s // 8) ** 2
        elif scope.startswith("res4") or scope.startswith("res5a") or scope.startswith("res5b"):
            px = (s // 16) ** 2
        else:  # res5c_branch2b
            px = (s // 8) ** 2
        macs += px * k * k * cin * cout
    macs += (s // 8) ** 2 * 4 * 256 * (63 + 128)  # two transposed convs
    macs += (s // 8) ** 2 * 128 * 84  # res5c_branch2c
    return 2 * macs
