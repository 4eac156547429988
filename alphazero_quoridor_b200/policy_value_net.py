"""Policy-value net: the reference's 5-block ResNet (policy_value_net.py:20-99) and its `PolicyValueNet`
wrapper (policy_value_net.py:109-200), kept in PyTorch -- it is the only dense contraction on the path
(BASELINE.json north_star) -- with a bf16 channels_last inference copy fed directly by the encode kernel.

State-dict keys match the reference (`conv1, bn1, res{1..5}.{conv1,bn1,conv2,bn2}, conv2, bn2, fc1, fc2,
conv3, bn3, fc3`), so `ckpt/<name>.pth` files interchange.

Deviations: batched inference runs BatchNorm in eval mode (the reference leaves the module in
training mode, so its batch-1 search normalises each leaf by its own statistics); `train_step` returns
Python floats (`loss.data[0]` at policy_value_net.py:192 raises on current PyTorch).
"""
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.optim as optim

from . import _lib

IN_PAD = 32      # input channels of the inference copy (26 planes zero-padded for channels_last bf16)


def set_learning_rate(optimizer, lr):
    for param_group in optimizer.param_groups:
        param_group['lr'] = lr


def conv3x3(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False)


class BasicBlock(nn.Module):
    """policy_value_net.py:20-48"""

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super(BasicBlock, self).__init__()
        self.conv1 = conv3x3(inplanes, planes, stride)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = conv3x3(planes, planes)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        residual = x
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        if self.downsample is not None:
            residual = self.downsample(x)
        out = out + residual
        return self.relu(out)


class policy_value_net(nn.Module):
    """policy_value_net.py:51-99: 26x9x9 -> (log-probs [B,140], value [B,1])."""

    def __init__(self, block=BasicBlock, inplanes=26, planes=64, stride=1):
        super(policy_value_net, self).__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.res1 = block(planes, planes)
        self.res2 = block(planes, planes)
        self.res3 = block(planes, planes)
        self.res4 = block(planes, planes)
        self.res5 = block(planes, planes)
        self.conv2 = nn.Conv2d(64, 4, kernel_size=3, stride=stride, padding=1, bias=False)   # value head
        self.bn2 = nn.BatchNorm2d(4)
        self.fc1 = nn.Linear(324, 128)
        self.fc2 = nn.Linear(128, 1)
        self.conv3 = nn.Conv2d(64, 2, kernel_size=3, stride=stride, padding=1, bias=False)   # policy head
        self.bn3 = nn.BatchNorm2d(2)
        self.fc3 = nn.Linear(162, 140)

    def forward(self, x):
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.res5(self.res4(self.res3(self.res2(self.res1(out)))))
        v = self.relu(self.bn2(self.conv2(out)))
        v = v.reshape(-1, 324)                       # NCHW flatten order, as .view(-1, 324) in the reference
        v = torch.tanh(self.fc2(self.fc1(v)))        # fc1 has no activation in the reference (:94-95)
        p = self.relu(self.bn3(self.conv3(out)))
        p = p.reshape(-1, 162)
        p = F.log_softmax(self.fc3(p), dim=1)
        return p, v


class PolicyValueNet(object):
    """policy_value_net.py:109-200"""

    def __init__(self, model_file=None, use_gpu=True, device=None, infer_dtype=torch.bfloat16, max_batch=8192):
        self.use_gpu = use_gpu
        self.l2_const = 1e-4
        if use_gpu:
            _lib.require_cuda()
            self.device = torch.device(device if device is not None else "cuda")
            if self.device.index is None:
                self.device = torch.device("cuda", torch.cuda.current_device())
        else:
            self.device = torch.device("cpu")
        self.policy_value_net = policy_value_net(BasicBlock, 26, 64).to(self.device)
        self.optimizer = optim.Adam(self.policy_value_net.parameters(), weight_decay=self.l2_const)
        if model_file:
            self.policy_value_net.load_state_dict(torch.load('ckpt/%s.pth' % model_file, map_location=self.device))
        self.infer_dtype = infer_dtype
        self.max_batch = max_batch
        self._infer = None
        self._in_buf = None
        self._env = None
        # the 12-launch inference forward (x chunks) replayed as ONE CUDA graph per batch size: at 8192-position chunks a
        # layer runs for ~80 us, so launch gaps are a visible share of the AlphaZero path
        self.use_cuda_graph = bool(use_gpu)
        self._graphs = {}

    # ---- batched device path (the hot one) ----
    @staticmethod
    def _fold(conv, bn, pad_in=None, dtype=torch.bfloat16):
        """Conv + eval-mode BatchNorm -> one conv with bias (folded in fp32, then cast)."""
        w = conv.weight.detach().float()
        g = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
        b = bn.bias.detach().float() - bn.running_mean.detach().float() * g
        w = w * g.view(-1, 1, 1, 1)
        if pad_in:
            wp = torch.zeros(w.shape[0], pad_in, w.shape[2], w.shape[3], device=w.device)
            wp[:, :w.shape[1]] = w
            w = wp
        return w.to(dtype).contiguous(memory_format=torch.channels_last), b.to(dtype).contiguous()

    def sync_inference_weights(self):
        """(Re)build the inference copy used by the search; call after training updates.
        Eval-mode BatchNorm is folded into the convolutions and every layer becomes one cuDNN fused
        conv+bias(+residual)+ReLU call on channels_last bf16 tensors (13 convs -> 12 launches: the two 3x3 head
        convolutions share one); 3.9x faster than running the nn.Module in bf16 on B200 and closer to fp32."""
        m = self.policy_value_net
        dt = self.infer_dtype
        W = {"c1": self._fold(m.conv1, m.bn1, IN_PAD, dt)}
        for i in range(1, 6):
            blk = getattr(m, "res%d" % i)
            W["r%da" % i] = self._fold(blk.conv1, blk.bn1, None, dt)
            W["r%db" % i] = self._fold(blk.conv2, blk.bn2, None, dt)
        wv, bv = self._fold(m.conv2, m.bn2, None, dt)
        wp, bp = self._fold(m.conv3, m.bn3, None, dt)
        W["head"] = (torch.cat([wv, wp], 0).contiguous(memory_format=torch.channels_last), torch.cat([bv, bp], 0))
        W["fc"] = [(l.weight.detach().to(dt).contiguous(), l.bias.detach().to(dt).contiguous()) for l in (m.fc1, m.fc2, m.fc3)]
        self._infer = W
        self._graphs = {}                       # captured graphs hold the old weight tensors
        return W

    def _infer_forward(self, x):
        W = self._infer
        S, P, D = [1, 1], [1, 1], [1, 1]
        out = torch.cudnn_convolution_relu(x, W["c1"][0], W["c1"][1], S, P, D, 1)
        for i in range(1, 6):
            a, b = W["r%da" % i], W["r%db" % i]
            t = torch.cudnn_convolution_relu(out, a[0], a[1], S, P, D, 1)
            out = torch.cudnn_convolution_add_relu(t, b[0], out, 1.0, b[1], S, P, D, 1)       # BasicBlock, :31-48
        h = torch.cudnn_convolution_relu(out, W["head"][0], W["head"][1], S, P, D, 1)         # value (4) + policy (2) planes
        v = h[:, :4].reshape(-1, 324)                                                          # NCHW flatten order (:92,:99)
        p = h[:, 4:].reshape(-1, 162)
        (w1, b1), (w2, b2), (w3, b3) = W["fc"]
        v = torch.tanh(F.linear(F.linear(v, w1, b1), w2, b2).float())
        p = F.log_softmax(F.linear(p, w3, b3).float(), dim=1)
        return p, v

    def _forward_chunks(self, x):
        m = x.shape[0]
        if m <= self.max_batch:
            logp, v = self._infer_forward(x)
        else:
            # chunks whose activations (64 ch x 81 x 2 B = 10 KB per position) stay inside the 126 MB L2
            outs = [self._infer_forward(x[i:i + self.max_batch]) for i in range(0, m, self.max_batch)]
            logp, v = torch.cat([o[0] for o in outs], 0), torch.cat([o[1] for o in outs], 0)
        return logp.exp().contiguous(), v.reshape(-1).contiguous()

    def _graph_for(self, m):
        """Capture the inference forward on the first m rows of the (static) input buffer; returns (graph, probs, value)
        or None when capture is unavailable (the eager path is used then -- still on the GPU)."""
        ent = self._graphs.get(m)
        if ent is not None or not self.use_cuda_graph:
            return ent
        x = self._in_buf[:m]
        try:
            cur = torch.cuda.current_stream(self.device)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side), torch.no_grad():
                for _ in range(2):
                    self._forward_chunks(x)
            cur.wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(graph):
                probs, value = self._forward_chunks(x)
            ent = (graph, probs, value)
        except Exception as ex:                 # keep running eagerly, say so once
            import warnings
            warnings.warn("CUDA graph capture of the inference forward failed (%r); running it eagerly" % (ex,))
            self.use_cuda_graph = False
            ent = None
        self._graphs[m] = ent
        return ent

    def evaluate_states(self, states):
        """qz_state rows int64 [m,3] (CUDA) -> (probs float32 [m,140], value float32 [m]), all on the device.
        The encode kernel writes the leaves straight into the net's channels_last bf16 input.  The returned tensors
        are reused by the next call with the same m (consume them in stream order, as the search does)."""
        if self._infer is None:
            self.sync_inference_weights()
        m = states.shape[0]
        if self._in_buf is None or self._in_buf.shape[0] < m:
            self._in_buf = torch.empty((max(m, 1), IN_PAD, 9, 9), dtype=self.infer_dtype, device=self.device,
                                       memory_format=torch.channels_last)
            self._graphs = {}                   # graphs captured on the old buffer are void
        x = self._in_buf[:m]
        lib = _lib.load()
        with torch.cuda.device(self.device):
            _lib.check(lib.qz_env_encode(_lib.ptr(states), _lib.c_void_p_of(x), _lib.DTYPE_CODE[self.infer_dtype],
                                         _lib.LAYOUT_NHWC, IN_PAD, m, _lib.stream_ptr(self.device)), "qz_env_encode")
            ent = self._graph_for(m) if m > 0 else None
            if ent is not None:
                ent[0].replay()
                return ent[1], ent[2]
            with torch.no_grad():
                return self._forward_chunks(x)

    # ---- reference API ----
    def _f32(self, x):
        """numpy arrays / lists (the reference's inputs) or tensors already on the device -> float32 on self.device."""
        if torch.is_tensor(x):
            return x.to(device=self.device, dtype=torch.float32)
        return torch.as_tensor(np.asarray(x), dtype=torch.float32, device=self.device)

    def policy_value(self, state_batch):
        """policy_value_net.py:127-143: [B,26,9,9] array -> (act_probs [B,140], value [B,1]) numpy."""
        x = self._f32(state_batch)
        with torch.no_grad():
            log_act_probs, value = self.policy_value_net(x)
        return np.exp(log_act_probs.cpu().numpy()), value.cpu().numpy()

    def policy_value_fn(self, game):
        """policy_value_net.py:145-164: game -> (zip(legal, probs[legal]) un-renormalised, value)."""
        legal_positions = game.actions()
        if self.use_gpu and hasattr(game, "packed"):
            st = torch.tensor([game.packed()], dtype=torch.int64, device=self.device)
            probs, value = self.evaluate_states(st)
            act_probs = probs[0].cpu().numpy()
            v = value[0].cpu()
        else:
            current_state = np.ascontiguousarray(game.state()).reshape([1, 26, 9, 9])
            with torch.no_grad():
                log_act_probs, value = self.policy_value_net(torch.from_numpy(current_state).float().to(self.device))
            act_probs = np.exp(log_act_probs.cpu().numpy().flatten())
            v = value[0][0].cpu()
        return zip(legal_positions, act_probs[legal_positions]), v

    def train_step(self, state_batch, mcts_probs, winner_batch, lr, grad_hook=None):
        """policy_value_net.py:166-192: loss = (z - v)^2 - pi^T log p (+ weight decay via Adam)."""
        state_batch, mcts_probs, winner_batch = self._f32(state_batch), self._f32(mcts_probs), self._f32(winner_batch)
        self.policy_value_net.train()
        self.optimizer.zero_grad()
        set_learning_rate(self.optimizer, lr)
        log_act_probs, value = self.policy_value_net(state_batch)
        value_loss = F.mse_loss(value.view(-1), winner_batch)
        policy_loss = -torch.mean(torch.sum(mcts_probs * log_act_probs, 1))
        loss = value_loss + policy_loss
        loss.backward()
        if grad_hook is not None:
            grad_hook()                         # e.g. NCCL all-reduce of the gradients (train.TrainPipeline)
        self.optimizer.step()
        entropy = -torch.mean(torch.sum(torch.exp(log_act_probs) * log_act_probs, 1))
        self._infer = None                      # inference copy is stale now
        return loss.item(), entropy.item()

    def get_policy_param(self):
        return self.policy_value_net.state_dict()

    def save_model(self, model_file):
        os.makedirs('ckpt', exist_ok=True)
        torch.save(self.policy_value_net.state_dict(), 'ckpt/%s.pth' % (model_file))
