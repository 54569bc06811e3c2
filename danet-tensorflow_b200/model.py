"""
Model assembly -- the role of the reference's `Model` (main.py:61-399) without a graph:
`build()` instantiates the registered Encoder / Estimator / Separator plugins,
`train_forward` / `valid_forward` / `infer` run what one `sess.run` of the matching fetch
list computes (main.py:369-397), `separate` is the demo path wav -> wavs
(main.py:660-695 + app/utils.py:111-135).  Variables live in `self.params` under the
reference's TF names so checkpoints are interchangeable by name.
"""
import itertools

import numpy as np
import torch

from . import kernels as K
from . import shard
from . import modules as _modules   # noqa: F401  (registers the plugins, like app/__init__.py:1-5)
from . import ozers as _ozers       # noqa: F401
from .hparams import hparams


class Model(object):
    def __init__(self, name='danet', device='cuda:0', seed=1337):
        self.name = name
        self.device = torch.device(device)
        self.params = {}                  # reference variable name (without 'global/') -> tensor
        self._rs = np.random.RandomState(seed)
        self.encoder = None
        self.estimator = None
        self.infer_estimator = None
        self.separator = None
        self.s_states_di = {}             # main.py:110-121 -- zero state, never assigned
        self._stagger_pending = False
        self._stagger_event = None
        self._streams = []
        self._twins = {}
        self._graphs = {}
        self._packed = {}                 # weights pre-split for the tensor cores (dropped whenever they change)
        self._packed_ready = False        # True once every entry exists (built on ONE stream, see _prepare_packed)
        self._last_split = None           # (hidden sequence, its split copy, rows time-major?) handed from layer to layer
        self._group_grad_bufs = {}        # training slice g >= 1 -> (flat gradient buffer, per-variable views)
        self._train_graphs, self._train_warm = {}, {}     # captured forward + backward per input shape
        self._in_train_group, self._train_group = False, 0
        self._tape = None                 # training: saved activations per recurrent layer
        self._flat = None                 # training: (param, grad, adam m, adam v) flat buffers + views
        self._buckets = None              # training: per-layer all-reduce of slices of the flat gradient buffer
        self._ar_events = []
        self._vars_ready = False          # every variable exists (created in the reference's order)
        self._flag_pool = None            # [tensor [FLAG_SETS, 64] int32 cleared for the running group, sets handed out]
        self._fuse_anchor = None          # inference: anchors handed to the encoder's projection (fused estimator sums)
        self._fused_attrs = None          # (embedding, attractors) the fused projection produced
        self.step_count = 0

    # ---------------------------------------------------------------- variables
    def get_variable(self, name, shape, init):
        """tf.get_variable with reuse: created on first use, in call order, from one RandomState
        stream (same stream as the golden fixtures' weights)."""
        v = self.params.get(name)
        if v is None:
            a = np.ascontiguousarray(np.asarray(init(self._rs, list(shape)), dtype=np.float32))
            if list(a.shape) != list(shape):
                raise ValueError('initialiser for %s returned %s, wanted %s' % (name, a.shape, shape))
            v = torch.from_numpy(a).to(self.device)
            self.params[name] = v
        elif list(v.shape) != list(shape):
            raise ValueError('variable %s has shape %s, requested %s' % (name, tuple(v.shape), shape))
        return v

    def _invalidate(self):
        """the weights changed (or moved): drop everything derived from them -- the pre-split tensor-core images AND the
        captured CUDA graphs, which hold raw pointers to those images and to the variables themselves"""
        self._packed = {}
        self._packed_ready = False
        self._graphs = {}
        self._last_split = None

    def load_params(self, params):
        """Take weights by reference name (numpy arrays or tensors); the role of main.py:201-206.  A variable the model
        already holds must keep its shape (tf.train.Saver.restore raises on a mismatch)."""
        self._invalidate()
        for k, v in params.items():
            t = torch.as_tensor(np.asarray(v) if not isinstance(v, torch.Tensor) else v)
            t = t.to(device=self.device, dtype=torch.float32).contiguous()
            if k in self.params and tuple(self.params[k].shape) != tuple(t.shape):
                raise ValueError('variable %s has shape %s, the restored value %s'
                                 % (k, tuple(self.params[k].shape), tuple(t.shape)))
            if self._flat is not None and k in self.params:
                self.params[k].copy_(t)          # keep the flat-buffer views
            else:
                self.params[k] = t

    TF_SCOPE = 'global/'               # main.py:229: every trainable variable of the reference lives under this scope

    def save_params(self, path):
        """main.py:192-199 -- trainable variables only (no optimiser slots).  `*.pt`: a torch file of {name: tensor};
        anything else: a TensorFlow checkpoint bundle `<path>.index` + `<path>.data-00000-of-00001` under the
        reference's variable names (tf_bundle.py), the format its tf.train.Saver writes"""
        if path.endswith('.pt'):
            torch.save({k: v.cpu() for k, v in self.params.items()}, path)
            return path
        from . import tf_bundle
        return tf_bundle.write_bundle(path, {self.TF_SCOPE + k: v.detach().cpu().numpy() for k, v in self.params.items()})

    def load_params_file(self, path, strict=True):
        """main.py:201-206: restore from a torch file (`*.pt`) or a TensorFlow checkpoint prefix.  Variables the model
        does not have (optimiser slots, step counters) are ignored; returns the names that were loaded.  Like
        tf.train.Saver.restore (NotFoundError), a model variable the file does not hold is an error -- a checkpoint of
        another encoder / estimator configuration must not restore partially and report success (`strict=False`: only
        what matches by name is loaded)."""
        if path.endswith('.pt'):
            params = torch.load(path)
        else:
            from . import tf_bundle
            prefix = path[:-6] if path.endswith('.index') else path
            n = len(self.TF_SCOPE)
            params = {k[n:]: v for k, v in tf_bundle.read_bundle(prefix).items() if k.startswith(self.TF_SCOPE)}
        if self.params:
            params = {k: v for k, v in params.items() if k in self.params}
            missing = sorted(set(self.params) - set(params))
            if missing and strict:
                raise KeyError('checkpoint %s does not hold %d of the model\'s %d variables (first: %s)'
                               % (path, len(missing), len(self.params), missing[0]))
        self.load_params(params)
        return sorted(params)

    def parameter_count(self):
        return sum(int(v.numel()) for v in self.params.values())

    # ---------------------------------------------------------------- recurrent layers
    def _lstm_vars(self, name, idim, hdim, w_init, b_init):
        W = self.get_variable(name + '/LSTM/linear/W', [idim + hdim, 4 * hdim], w_init)
        Bv = self.get_variable(name + '/LSTM/linear/B', [4 * hdim], b_init)
        return W, Bv

    def lyr_lstm(self, name, s_x, hdim, axis=-1, t_axis=0, op_linear=None, w_init=None, b_init=None,
                 reverse=False):
        """main.py:76-132: one direction over [B,T,I] from zero state -> [B,T,hdim].
        The layer is `hoisted input GEMM for all t` + `persistent recurrent kernel`."""
        B, T, I = s_x.shape
        W, Bv = self._lstm_vars(name, I, hdim, w_init, b_init)
        x2 = s_x.reshape(B * T, I)
        pre = K.linear(x2, W, Bv, time_major_T=T, k_rows=I).view(1, T, B, 4 * hdim)
        if reverse:
            raise NotImplementedError('single reversed direction: use lyr_bilstm')
        wide = K.DEFAULT_BACKEND == 1 and K.TC_LSTM_MAX_H < hdim <= K.TC_WIDE_MAX_H
        if self._tape is not None:
            # training forward of a wide layer (lstm-orig): the wide tcgen05 kernel keeps cell states and gates as the
            # exact-fp32 kernel does (the weights change every step, so the image is packed by the library on the fly)
            out, cell = K.lstm_seq(pre, [W], I, T, B, hdim, keep_cell=True, keep_gates=True,
                                   backend=2 if wide and self.train_recurrent_fp16() else None)
            self._tape.append(dict(name=name, x=s_x, gates=pre, cell=cell, out=out, hdim=hdim))
            return out
        if self.RECURRENT_FP16 and wide:
            # lstm-orig (H = 600) in inference: the wide tcgen05 kernel on a cached weight image (fp16 state, C-ABI backend 2)
            wh_packed = self._packed.get(name + '/wh')
            if wh_packed is None:
                wh_packed = self._packed[name + '/wh'] = K.lstm_pack_wh([W], I, hdim)
            return K.lstm_seq(pre, [W], I, T, B, hdim, backend=2, wh_packed=wh_packed)
        return K.lstm_seq(pre, [W], I, T, B, hdim)

    def lyr_bilstm(self, name, s_x, hdim, w_init=None, b_init=None):
        """app/modules.py:120-137 in one launch: [B,T,I] -> [B,T,2*hdim] (fwd | bwd un-reversed)"""
        B, T, I = s_x.shape
        Wf, Bf = self._lstm_vars(name + '_fwd', I, hdim, w_init, b_init)
        Wb, Bb = self._lstm_vars(name + '_bwd', I, hdim, w_init, b_init)
        if self._tape is not None:
            return self._lyr_bilstm_train(name, s_x, hdim, (Wf, Bf, Wb, Bb))
        if self.USE_PACKED and K.DEFAULT_BACKEND == 1 and hdim <= K.TC_LSTM_MAX_H:
            return self._lyr_bilstm_packed(name, s_x, hdim, (Wf, Bf, Wb, Bb))
        x2 = s_x.reshape(B * T, I)
        pre = torch.empty((2, T, B, 4 * hdim), dtype=torch.float32, device=s_x.device)
        K.linear(x2, Wf, Bf, time_major_T=T, k_rows=I, out=pre[0].view(T * B, 4 * hdim))
        K.linear(x2, Wb, Bb, time_major_T=T, k_rows=I, out=pre[1].view(T * B, 4 * hdim))
        if self._stagger_pending:          # see separate(): the next stream group may start now
            self._stagger_pending = False
            self._stagger_event = torch.cuda.current_stream().record_event()
        K.stamp('%s gemm' % name)
        out = K.lstm_seq(pre, [Wf, Wb], I, T, B, hdim)
        K.stamp('%s lstm' % name)
        return out

    # ---------------------------------------------------------------- inference fast path
    USE_PACKED = True          # one product per layer on cached pre-split weights + operands emitted by the LSTM
    # Inference carries h_{t-1} into the recurrent product as ONE fp16 value (C-ABI backend 2) instead of a bf16 hi/lo
    # pair: half the DSMEM bytes per step, ~1e-4 max-norm deviation of the embedding (tools/precision_study.py) against
    # the 1e-3 parity gate.  False = bf16x3 everywhere (~1e-5).  Training always uses bf16x3.
    RECURRENT_FP16 = True

    def recurrent_backend(self):
        """C-ABI backend of the inference recurrence: 2 (fp16 state) unless switched off"""
        return 2 if self.RECURRENT_FP16 else 1
    USE_CENTER_FOLD = True     # output projection with the centring folded into its epilogue
    PIPELINE_INPUT_GEMM = True # the recurrence starts on the first finished tiles of its input projections
    TIME_MAJOR_HANDOVER = True # ... whose A operand has time-major rows, so that 2 finished row tiles are enough to start
    PROGRAMMATIC_LSTM_LAUNCH = True    # layer l+1's recurrence queued behind layer l's as a programmatic dependent (PDL)
    _pdl_prev = None           # (priority stream, output) of the last pipelined recurrence queued
    _flags_precleared = False
    _emit_time_major = False   # set by the encoder around a layer whose output feeds another pipelined recurrent layer
    # training forward: carry h into the recurrent product as fp16 as inference does (0.57 -> 0.47 ms per layer at cfg 2,
    # step 9.38 -> 9.14 ms).  None = the encoder decides (its TRAIN_FP16_STATE_OK): on for the purely recurrent encoders --
    # against float64 autograd on the oracle the bilstm-orig gradients move only from 2.9e-5 to 4.2e-5 of their variable's
    # largest entry (tools/train_precision.py) -- and off for conv-bilstm-v1, whose reference-generated gradient fixture
    # fails its 1e-3 gate with it (first convolution's kernel: 3.3e-3 -- the max-pools turn the 1e-4 forward deviation into
    # routing changes).  True / False force it for A/B runs.
    TRAIN_RECURRENT_FP16 = None

    def train_recurrent_fp16(self):
        if self.TRAIN_RECURRENT_FP16 is not None:
            return bool(self.TRAIN_RECURRENT_FP16)
        return bool(getattr(self.encoder, 'TRAIN_FP16_STATE_OK', False))

    FLAG_SETS = 8              # completion-flag sets cleared by ONE launch per group and step (one set per recurrent layer)

    def _next_flags(self, device):
        """64 cleared completion flags for one pipelined product.  `separate` clears FLAG_SETS sets with a single launch
        when a group starts (off the per-layer critical path) and the layers take them in turn; outside that (or when
        the pool is used up) a set is cleared on the spot."""
        pool = self._flag_pool
        self._flags_precleared = pool is not None and pool[1] < self.FLAG_SETS
        if self._flags_precleared:
            pool[1] += 1
            return pool[0][pool[1] - 1]
        return K.pipeline_flags(device)

    def _lyr_bilstm_packed(self, name, s_x, hdim, weights):
        """Same arithmetic as lyr_bilstm with the operand traffic trimmed: the two directions' input weights are
        split to bf16 hi/lo ONCE and kept side by side (one product, N = 8H, instead of two), and the layer input
        arrives already split from the previous layer's recurrent kernel, so the dense layer is a single launch."""
        Wf, Bf, Wb, Bb = weights
        B, T, I = s_x.shape
        ent = self._packed.get(name)
        if ent is None:
            N = 8 * hdim
            w2 = K.split_operand(Wf[:I], True, rows_total=N, row0=0)
            K.split_operand(Wb[:I], True, out=w2, rows_total=N, row0=4 * hdim)
            ent = (w2, torch.cat([Bf, Bb]), K.lstm_pack_wh([Wf, Wb], I, hdim))
            self._packed[name] = ent
        w2, bias2, wh_packed = ent
        prev = self._last_split
        backend = self.recurrent_backend()
        pipelined = (self.PIPELINE_INPUT_GEMM and self.LSTM_PRIORITY_STREAM and B * T <= K.PIPELINE_MAX_ROWS
                     and B <= self.PIPELINE_GROUP and wh_packed is not None)
        # rows of the product's A operand in TIME-major order (t*B + b): the first / last row tile then serve the first /
        # last 16 steps of every utterance and the scans start after 2 tiles of the product instead of ~9 of its 32
        if prev is not None and prev[0] is s_x:
            a2, a_tm = prev[1], prev[2]
        elif pipelined and self.TIME_MAJOR_HANDOVER:
            a2, a_tm = K.split_operand_time_major(s_x.reshape(B * T, I), T), True
        else:
            a2, a_tm = K.split_operand(s_x.reshape(B * T, I), False), False
        if pipelined:
            # The recurrence does not wait for the whole product: the GEMM issues its row tiles in the order the two scans
            # consume them and publishes each through a flag; the recurrent kernel, launched from its high-priority stream
            # as soon as the product has been QUEUED, spins on the flag of the tile it is about to read
            # (danet_gemm_split_pipelined / danet_lstm_seq_fwd_pipelined).  ~40 of the product's ~50 us leave the group's
            # critical path, four times per step.
            cur = torch.cuda.current_stream()
            hp_stream = self._priority_twin(cur)
            flags = self._next_flags(s_x.device)
            queued = cur.record_event()
            if self.LSTM_HEADSTART_US > 0:
                # give the recurrence's 10-CTA clusters first pick of the SMs: launched at the same moment, the product's
                # single CTAs fill SMs one by one and a cluster never finds ten free ones in a GPC until the product drains
                torch.cuda._sleep(int(self.LSTM_HEADSTART_US * 1840))
            pre, need = K.gemm_split_pipelined(a2, w2, B * T, 8 * hdim, I, T, flags, bias=bias2, rows_tm=a_tm)
            pre = pre.view(T, B, 2, 4 * hdim)
            if self._stagger_pending:
                self._stagger_pending = False
                # STAGGER_US > 0: the next group is released that long after this product STARTED instead of when it ends
                self._stagger_event = queued if self.STAGGER_US > 0 else cur.record_event()
            K.stamp('%s gemm' % name)
            # Programmatic dependent launch: when this layer's input IS the previous pipelined layer's output, that layer's
            # recurrence is the last thing queued on hp_stream, and this one needs nothing else that is not guarded by the
            # flags (its flag set was cleared when the group started): it is queued directly behind it WITHOUT waiting for
            # this stream, as a programmatic dependent -- its clusters take over the SMs the previous recurrence frees as it
            # frees them (queued normally they find the product's CTAs there first: measured 23 us from one layer's exit to
            # the next one's entry) and run their prologue while it drains.
            pdl = bool(self.PROGRAMMATIC_LSTM_LAUNCH and self._flags_precleared and self._pdl_prev is not None
                       and self._pdl_prev[0] is hp_stream and self._pdl_prev[1] is s_x)
            if not pdl:
                hp_stream.wait_event(queued)
            with torch.cuda.stream(hp_stream):
                # only a layer that feeds another recurrent layer may emit time-major rows (the output projection cuts
                # its row tiles per utterance): the encoder says so through _emit_time_major
                emit_tm = bool(self.TIME_MAJOR_HANDOVER and self._emit_time_major)
                out, out_split = K.lstm_seq_pipelined(pre, [Wf, Wb], I, T, B, hdim, flags, need, backend=backend,
                                                      wh_packed=wh_packed, pre_tm=a_tm, split_tm=emit_tm, programmatic=pdl)
            cur.wait_stream(hp_stream)
            K.stamp('%s lstm' % name)
            self._last_split = (out, out_split, emit_tm)
            self._pdl_prev = (hp_stream, out)
            return out
        pre = K.gemm_split(a2, w2, B * T, 8 * hdim, I, bias=bias2, out_perm_T=0 if a_tm else T).view(T, B, 2, 4 * hdim)
        if self._stagger_pending:          # see separate(): the next stream group may start now
            self._stagger_pending = False
            self._stagger_event = torch.cuda.current_stream().record_event()
        K.stamp('%s gemm' % name)
        if self.LSTM_PRIORITY_STREAM:
            # the recurrence is the critical path and needs whole SMs (one cluster = 10 of them): launch it from a
            # high-priority twin of this group's stream so its CTAs are placed before other groups' dense tiles
            cur = torch.cuda.current_stream()
            hp_stream = self._priority_twin(cur)
            hp_stream.wait_stream(cur)
            with torch.cuda.stream(hp_stream):
                out, out_split = K.lstm_seq(pre, [Wf, Wb], I, T, B, hdim, interleaved=True, want_split=True,
                                            wh_packed=wh_packed, backend=backend)
            cur.wait_stream(hp_stream)
        else:
            out, out_split = K.lstm_seq(pre, [Wf, Wb], I, T, B, hdim, interleaved=True, want_split=True,
                                        wh_packed=wh_packed, backend=backend)
        K.stamp('%s lstm' % name)
        self._last_split = (out, out_split, False)
        return out

    def centered_projection(self, name, x, W):
        """(x - mean_b(x)) @ W for x [B,T,K] (app/modules.py:244-255) as ONE product on the operand the last
        recurrent layer already emitted: (x - mu) W = x W - mu colsum(W), the rank-1 term applied in the epilogue.
        Returns None when the fast path does not apply."""
        prev = self._last_split
        if (not self.USE_CENTER_FOLD or self._tape is not None or K.DEFAULT_BACKEND != 1 or prev is None or prev[0] is not x
                or prev[2]):                   # prev[2]: the operand's rows are time-major (never for the last layer)
            return None
        ent = self._packed.get(name)
        if ent is None:
            ent = self._packed[name] = (K.split_operand(W, True), K.colsum(W))
        w2, col_s = ent
        B, T, Kd = x.shape
        if self._fuse_anchor is not None:
            # the anchor estimator's sums are taken in this product's epilogue (danet_proj_anchor_fwd);
            # AnchoredEstimator.__call__ picks the attractors up instead of launching its own pass over the embedding
            E = hparams.EMBED_SIZE
            embed, attrs = K.proj_anchor(prev[1], w2, B, T, W.shape[1] // E, E, Kd, self._fuse_anchor, row_mu=K.mean(x),
                                         col_s=col_s)
            self._fused_attrs = (embed, attrs)
            return embed.view(B * T, W.shape[1])
        return K.gemm_split(prev[1], w2, B * T, W.shape[1], Kd, row_mu=K.mean(x), col_s=col_s, rows_per_mu=T)

    USE_FUSED_PROJ_ANCHOR = True     # inference: projection + anchor-estimator sums in one kernel when the shapes allow

    def _anchors_for_fusion(self):
        """the infer estimator's anchors when the fused projection applies (exactly the `anchor` estimator, two sources,
        E = 20, at most 6 anchors, tensor-core backend, variables already created), else None"""
        est = self.infer_estimator
        if (not self.USE_FUSED_PROJ_ANCHOR or self._tape is not None or K.DEFAULT_BACKEND != 1 or hparams.DEBUG
                or type(est) is not _modules.AnchoredEstimator or hparams.MAX_N_SIGNAL != 2
                or hparams.EMBED_SIZE != K.PROJ_ANCHOR_E or hparams.NUM_ANCHOR > 6 or not self.USE_CENTER_FOLD):
            return None
        self._ensure_variables()
        return self.params.get('%s/anchors' % est.name)

    def _ensure_variables(self):
        """very first inference call: create every variable in the reference's order (encoder, train estimator, infer
        estimator: main.py:210-270) with the 4-frame dry run of reset(), so that the first call already takes the same
        (fused) path as every later one"""
        if self._vars_ready:
            return
        pending, self._stagger_pending = self._stagger_pending, False
        fuse, self._fuse_anchor = self._fuse_anchor, None
        self._vars_ready = True
        self.reset()
        self._stagger_pending, self._fuse_anchor = pending, fuse

    def dense(self, name, x2, W, bias=None):
        """x2 [M,K] @ W (+ bias) with the weight operand split once and cached (inference)"""
        if self._tape is not None or K.DEFAULT_BACKEND != 1:
            return K.linear(x2, W, bias)
        w2 = self._packed.get(name)
        if w2 is None:
            w2 = self._packed[name] = K.split_operand(W, True)
        return K.gemm_split(K.split_operand(x2, False), w2, x2.shape[0], W.shape[1], W.shape[0], bias=bias)

    # ---------------------------------------------------------------- training step (row a16)
    def _lyr_bilstm_train(self, name, s_x, hdim, weights):
        """lyr_bilstm keeping what the backward needs (gates in place of the pre-activations, cell states).  The layer
        input is split to bf16 hi/lo once for both directions -- or not at all when the previous layer's recurrent kernel
        already emitted it -- and the weights once per step (they change every step)."""
        Wf, Bf, Wb, Bb = weights
        B, T, I = s_x.shape
        x2 = s_x.reshape(B * T, I)
        pre = torch.empty((2, T, B, 4 * hdim), dtype=torch.float32, device=s_x.device)
        tc = K.DEFAULT_BACKEND == 1 and hdim <= K.TC_LSTM_MAX_H
        if tc:
            prev = self._last_split
            a2 = prev[1] if prev is not None and prev[0] is s_x else K.split_operand(x2, False)
            for d, (W, Bv) in enumerate(((Wf, Bf), (Wb, Bb))):
                K.gemm_split(a2, K.split_operand(W[:I], True), B * T, 4 * hdim, I, bias=Bv, out_perm_T=T,
                             out=pre[d].view(T * B, 4 * hdim))
            if self._stagger_pending:      # train_forward_backward in stream groups: the next slice may start now
                self._stagger_pending = False
                self._stagger_event = torch.cuda.current_stream().record_event()
            out, cell, out_split = K.lstm_seq(pre, [Wf, Wb], I, T, B, hdim, keep_cell=True, keep_gates=True, want_split=True,
                                              backend=2 if self.train_recurrent_fp16() else None)
            self._last_split = (out, out_split, False)
        else:
            K.linear(x2, Wf, Bf, time_major_T=T, k_rows=I, out=pre[0].view(T * B, 4 * hdim))
            K.linear(x2, Wb, Bb, time_major_T=T, k_rows=I, out=pre[1].view(T * B, 4 * hdim))
            out, cell = K.lstm_seq(pre, [Wf, Wb], I, T, B, hdim, keep_cell=True, keep_gates=True)
        self._tape.append(dict(name=name, x=s_x, gates=pre, cell=cell, out=out, hdim=hdim))
        return out

    def flatten_params(self):
        """Re-home every variable in ONE flat buffer (views keep the reference names), with matching flat
        gradient and Adam-moment buffers: the gradient all-reduce and the clip+Adam update are then one
        call each over ~9 M floats."""
        names = list(self.params)
        offs, total = shard.flat_layout({k: self.params[k].numel() for k in names})   # 256-byte aligned views
        self._invalidate()
        flat = torch.zeros(total, dtype=torch.float32, device=self.device)
        grad = torch.zeros_like(flat)
        for k in names:
            v = self.params[k]
            view = flat[offs[k]:offs[k] + v.numel()].view(v.shape)
            view.copy_(v)
            self.params[k] = view
        self._flat = dict(param=flat, grad=grad, m=torch.zeros_like(flat), v=torch.zeros_like(flat), offs=offs)
        self.grads = {k: grad[offs[k]:offs[k] + self.params[k].numel()].view(self.params[k].shape) for k in names}
        self._buckets = shard.GradientBuckets(grad, {k: (offs[k], offs[k] + self.params[k].numel()) for k in names})
        self._train_graphs, self._train_warm, self._group_grad_bufs = {}, {}, {}     # they point into the old buffers
        return self._flat

    BUCKETED_ALLREDUCE = True          # N > 1: all-reduce each layer's gradients as soon as they are queued (under BPTT)
    TIME_ALLREDUCE = False             # bench: CUDA events around the part of the exchange the step has to wait for

    def grads_ready(self, names):
        """called by an encoder's backward on the stream that produced the gradients of `names`: their slice of the flat
        buffer may be exchanged now (no-op on one rank)"""
        if self.BUCKETED_ALLREDUCE and self._flat is not None and not self._in_train_group:
            self._buckets.reduce(names)

    # Training step in stream groups (the inference schedule applied to training): the batch is cut into TRAIN_GROUPS
    # slices that run forward + backward on their own streams, one dense layer apart, so that one slice's products fill
    # the SMs the other slices' latency-bound recurrences (20 SMs per 8 utterances and direction) leave idle.  Every slice
    # writes its own flat gradient buffer; the step's gradient is their batch-weighted sum (the loss is a batch mean,
    # app/ops.py:406-431), exact up to summation order.  On N > 1 ranks the exchange then runs once after the sum instead of
    # in per-layer buckets under the backward pass.  0 / 1 = the whole batch in one pass.  Measured at cfg 2 (B = 32, graph
    # replay, tools/time_train_groups.py): 9.34 ms in one pass, 8.88 ms in two slices, 10.0 ms in four -- unlike inference
    # the step is not latency-bound once sliced: a slice of 8 alone takes 6.35 ms (16: 7.25 ms), but the dense / streaming
    # work of a step (3.4 ms of the whole machine) has only the 68 SMs the recurrences leave.
    TRAIN_GROUPS = 2
    TRAIN_GROUP_MIN = 8                # utterances per slice at least (one recurrent cluster's batch tile)

    def _train_group_count(self, B):
        n = int(self.TRAIN_GROUPS or 1)
        return max(1, min(n, B // max(1, int(self.TRAIN_GROUP_MIN))))

    def _group_grads(self, g):
        """flat gradient buffer + per-variable views of training slice g >= 1 (slice 0 writes the model's own)"""
        ent = self._group_grad_bufs.get(g)
        if ent is None or ent[0].numel() != self._flat['grad'].numel():
            buf = torch.zeros_like(self._flat['grad'])
            offs = self._flat['offs']
            views = {k: buf[offs[k]:offs[k] + v.numel()].view(v.shape) for k, v in self.params.items()}
            ent = self._group_grad_bufs[g] = (buf, views)
        return ent

    def train_forward_backward(self, src):
        """One forward + backward of the train loss (main.py:289, 357-358) on complex spectra src [B,C,T,F];
        fills self.grads (views of one flat buffer), returns dict(loss, snr).  No optimiser step."""
        if self.estimator is None:
            self.build()
        if self._flat is None:
            self.reset()
            self.flatten_params()
        est_name = hparams.TRAIN_ESTIMATOR_METHOD
        if est_name not in ('anchor', 'truth', 'truth-threshold', 'truth-weighted'):
            raise NotImplementedError('no backward for estimator %r' % est_name)
        src = src.to(self.device)
        B = src.shape[0]
        groups = self._train_group_count(B)
        if groups <= 1:
            return self._train_forward_backward_slice(src)
        self._ensure_variables()           # created on ONE stream, in the reference's order, before the slices fork
        main = torch.cuda.current_stream()
        fork = main.record_event()
        streams = self._side_streams(groups)
        own_grads, outs, sizes, prev = self.grads, [], [], None
        self._in_train_group = True        # grads_ready(): no per-layer exchange of a slice's partial gradient
        try:
            for g, st in enumerate(streams):
                lo, hi = shard.shard_bounds(B, g, groups)
                st.wait_event(fork)
                if prev is not None:
                    st.wait_event(prev)    # stagger: slice g starts when slice g-1 has queued its first dense layer
                with torch.cuda.stream(st):
                    self.grads = own_grads if g == 0 else self._group_grads(g)[1]
                    self._train_group = g
                    self._stagger_pending, self._stagger_event = True, None
                    outs.append(self._train_forward_backward_slice(src[lo:hi]))
                    prev = self._stagger_event
                    self._stagger_pending = False
                sizes.append(hi - lo)
        finally:
            self.grads, self._train_group, self._in_train_group = own_grads, 0, False
        for st in streams:
            main.wait_stream(st)
        flat = self._flat['grad']
        flat.mul_(sizes[0] / B)
        for g in range(1, groups):
            flat.add_(self._group_grads(g)[0], alpha=sizes[g] / B)
        w = [n / B for n in sizes]
        return dict(loss=sum(o['loss'] * wi for o, wi in zip(outs, w)), snr=sum(o['snr'] * wi for o, wi in zip(outs, w)),
                    perm_idx=torch.cat([o['perm_idx'] for o in outs]))

    def _train_forward_backward_slice(self, src):
        est_name = hparams.TRAIN_ESTIMATOR_METHOD
        B, Cn, T, F = src.shape
        E = hparams.EMBED_SIZE
        feats = K.mix_features(src)
        # ---- forward, recording what the backward needs
        enc = self.encoder
        self._tape = []
        try:
            embed = enc(feats['logmag'])
        finally:
            tape, self._tape = self._tape, None
        embed_flat = embed.view(B, T * F, E)
        if est_name == 'anchor':
            anchors = self.estimator.anchors()
            attrs, _, _, choice, den = K.attractor_anchor(embed, anchors, Cn, return_den=True)
        else:
            anchors, choice = None, None
            attrs, den = K.attractor_truth(embed, feats['src_pwr'], feats['mix_pwr'], est_name, return_den=True)
        sep = K.mask_cmul(embed_flat, attrs, feats['mix'], hparams.SEPARATOR_TYPE, want=('sep',))['sep']
        pit = K.pit_mse(src, sep)
        # ---- backward
        grads = self.grads
        g = K.head_bwd(embed, attrs, feats['mix'], src, pit['perm_idx'], hparams.SEPARATOR_TYPE, est_name,
                       src_pwr=feats['src_pwr'], mix_pwr=feats['mix_pwr'], anchors=anchors, choice=choice, den=den)
        if est_name == 'anchor':
            grads[self.estimator.name + '/anchors'].copy_(g['d_anchors'])
        enc.backward(g['d_embed'].view(B * T, F * E), tape)
        return dict(loss=pit['loss'][0], snr=pit['snr'].mean(), perm_idx=pit['perm_idx'])

    def all_reduce_grads(self):
        """the ONE collective of the system: sum of the flat gradient buffer over ranks, bucketed per layer
        in backward completion order (output projection + anchors, then L3 .. L0; shard.GradientBuckets) -- here only what
        is still in flight is waited for and whatever no bucket covered is exchanged.  The 1/world mean is folded into the
        clip + Adam kernel's grad_scale, i.e. the clip acts on the full-batch gradient (main.py:358-363)."""
        if not self.TIME_ALLREDUCE:
            return self._buckets.finish()
        e0 = torch.cuda.current_stream().record_event(torch.cuda.Event(enable_timing=True))
        scale = self._buckets.finish()
        e1 = torch.cuda.current_stream().record_event(torch.cuda.Event(enable_timing=True))
        self._ar_events.append((e0, e1))
        return scale

    def apply_gradients(self, grad_scale=1.):
        """main.py:359-363: clip_by_value(+-GRAD_CLIP_THRES) then the registered optimiser (app/ozers.py), one fused
        launch over the flat parameter buffer"""
        self.step_count += 1
        self._invalidate()                # captured graphs hold the old packed weights
        f = self._flat
        opt = hparams.get_optimizer()(hparams.LR)
        if opt['kind'] == 'adam':
            K.clip_adam(f['param'], f['grad'], f['m'], f['v'], self.step_count, lr=opt['lr'], clip=hparams.GRAD_CLIP_THRES,
                        beta1=opt['beta1'], beta2=opt['beta2'], eps=opt['eps'], grad_scale=grad_scale)
        else:
            K.clip_sgd(f['param'], f['grad'], opt['lr'], clip=hparams.GRAD_CLIP_THRES, grad_scale=grad_scale)

    # The forward + backward of a training step replayed from a CUDA graph captured once per input shape (after two eager
    # steps): with the batch in stream groups a step is ~1500 launches, and queued from Python they take longer than the
    # GPU needs (measured: 4 groups eager 14.4 ms per step against 9.5 ms for one pass).  The variables live in the flat
    # buffer and are updated in place, so the captured pointers stay valid; the exchange and the optimiser run outside.
    TRAIN_GRAPH = True

    def train_forward_backward_graphed(self, src):
        key = (tuple(src.shape), self._train_group_count(src.shape[0]), hparams.TRAIN_ESTIMATOR_METHOD,
               hparams.SEPARATOR_TYPE)
        ent = self._train_graphs.get(key)
        if ent is None:
            n = self._train_warm.get(key, 0)
            if n < 2 or self._flat is None:        # variables, flat buffers, allocator, kernel attributes: eagerly first
                self._train_warm[key] = n + 1
                return self.train_forward_backward(src)
            static_src = src.to(self.device).clone()
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self.train_forward_backward(static_src)
            ent = self._train_graphs[key] = (graph, static_src, out)
        graph, static_src, out = ent
        if src.data_ptr() != static_src.data_ptr():
            static_src.copy_(src, non_blocking=True)
        graph.replay()
        return out

    def train_step(self, src):
        """train fetches (main.py:369-375): forward, backward, gradient all-reduce, clip + optimiser"""
        out = self.train_forward_backward_graphed(src) if self.TRAIN_GRAPH else self.train_forward_backward(src)
        self.apply_gradients(self.all_reduce_grads())
        return out

    def set_learn_rate(self, lr):
        hparams.LR = float(lr)                                   # main.py:541-548

    def get_learn_rate(self):
        return hparams.LR

    # ---------------------------------------------------------------- assembly
    def build(self):
        """main.py:208-278: pick the plugins; variables appear on first call, in the reference's
        creation order (encoder, train_estimator, infer_estimator)."""
        self.encoder = hparams.get_encoder()(self, 'encoder')
        self.estimator = hparams.get_estimator(hparams.TRAIN_ESTIMATOR_METHOD)(self, 'train_estimator')
        if hparams.INFER_ESTIMATOR_METHOD != hparams.TRAIN_ESTIMATOR_METHOD:
            self.infer_estimator = hparams.get_estimator(hparams.INFER_ESTIMATOR_METHOD)(self, 'infer_estimator')
            if self.infer_estimator.USE_TRUTH:
                raise ValueError('the inference estimator must not use ground truth')   # main.py:266-267
        else:
            self.infer_estimator = self.estimator
        self.separator = hparams.get_separator(hparams.SEPARATOR_TYPE)(self, 'separator')
        return self

    def reset(self):
        """main.py:534-537 initialises variables; here: materialise them with a 1-frame dry run"""
        F, Cn = hparams.FEATURE_SIZE, hparams.MAX_N_SIGNAL
        src = torch.zeros((1, Cn, 4, F), dtype=torch.complex64, device=self.device)
        self.train_forward(src)

    def reset_state(self):
        """main.py:538-540 -- a no-op by SURVEY.md F7"""

    @staticmethod
    def _perm_table(Cn, device):
        return torch.tensor(list(itertools.permutations(range(Cn))), dtype=torch.int64, device=device)

    def _separate_spectra(self, estimator, feats, embed, src_pwr=None):
        B, T, F, E = embed.shape
        embed_flat = embed.view(B, T * F, E)
        attrs = estimator(embed, s_src_pwr=src_pwr, s_mix_pwr=feats['mix_pwr'], s_embed_flat=embed_flat)
        out = self.separator(feats['mix_pwr'], attrs, embed_flat, s_mixed_signals=feats['mix'],
                             want=('sep_pwr', 'sep', 'masks'))
        out['attrs'] = attrs
        return out

    def train_forward(self, src):
        """
        Forward half of the train / valid / debug fetches (main.py:369-397) for complex spectra
        src [B,C,T,F]: everything `Model.build` wires at main.py:233-337.
        """
        src = src.to(self.device)
        B, Cn, T, F = src.shape
        feats = K.mix_features(src)                                              # main.py:233-240
        embed = self.encoder(feats['logmag'])                                    # main.py:243
        tr = self._separate_spectra(self.estimator, feats, embed, feats['src_pwr'])      # :249-254, :271-284
        if self.infer_estimator is self.estimator:
            va = tr                                                              # main.py:259-260
        else:
            va = self._separate_spectra(self.infer_estimator, feats, embed)      # main.py:263-278
        pit = K.pit_mse(src, tr['sep'])                                          # main.py:289
        perms = self._perm_table(Cn, src.device)
        sel = perms[pit['perm_idx'].long()]                                      # main.py:293-306
        bidx = torch.arange(B, device=src.device).unsqueeze(1)
        output = tr['sep'][bidx, sel]
        pit_v = K.pit_mse(feats['src_pwr'], va['sep_pwr'])                       # main.py:312
        # valid SNR (main.py:336): complex error of the re-phased, PIT-aligned magnitudes
        sel_v = perms[pit_v['perm_idx'].long()]
        cross_v = K.pit_mse(src, va['sep'])
        noise = cross_v['cross'].gather(2, sel_v.unsqueeze(-1)).squeeze(-1).mean(1)
        sig = (pit['cross'].new_zeros(B) + self._signal_power(src))
        eps = hparams.EPS
        valid_snr = (4.342944819 * (torch.log(sig + eps) - torch.log(noise + eps))).mean()
        return dict(embed=embed, attrs=tr['attrs'], attrs_valid=va['attrs'], masks=tr['masks'],
                    masks_valid=va['masks'], sep_pwr=tr['sep_pwr'], output=output,
                    train_loss=pit['loss'][0], train_snr=pit['snr'].mean(),
                    valid_loss=pit_v['loss'][0], valid_snr=valid_snr, infer_signals=va['sep'],
                    perm_idx=pit['perm_idx'], perm_losses=pit['perm_losses'],
                    mix=feats['mix'], mix_pwr=feats['mix_pwr'], logmag=feats['logmag'])

    @staticmethod
    def _signal_power(src):
        # mean |s|^2 over (C,T,F): reuse the PIT kernel's signal-power output through snr/noise algebra
        # is not possible for a different permutation, so take it from the cross matrix of (src, 0)
        z = torch.zeros_like(src)
        c = K.pit_mse(src, z)['cross']          # cross[b,i,j] = mean |s_i|^2
        return c[:, :, 0].mean(1)

    USE_FUSED_K4 = True        # separate(): mask x mixture -> iSTFT as one kernel (the separated spectra stay on chip)

    def infer(self, mix, logmag=None, want='sep', wav_out=None):
        """infer fetches (main.py:384-385): complex mixture [B,T,F] -> separated spectra [B,C,T,F]
        (want='wav': separated waveforms [B,C,64*T] through the fused back end)"""
        B, T, F = mix.shape
        if logmag is None:
            feats = K.mix_features(mix.view(B, 1, T, F), want=('mix_pwr', 'logmag'))
            logmag, mix_pwr = feats['logmag'], feats['mix_pwr']
        else:
            mix_pwr = None
        self._fused_attrs = None
        self._fuse_anchor = self._anchors_for_fusion()
        try:
            embed = self.encoder(logmag)
        finally:
            self._fuse_anchor = None
        K.stamp('proj')
        embed_flat = embed.view(B, T * F, -1)
        attrs = self.infer_estimator(embed, s_embed_flat=embed_flat)
        K.stamp('attractor')
        if want == 'wav':
            return self.separator(mix_pwr, attrs, embed_flat, s_mixed_signals=mix, want=('wav',), s_wav_out=wav_out)['wav']
        out = self.separator(mix_pwr, attrs, embed_flat, s_mixed_signals=mix, want=(want,))
        return out[want]

    def _prepare_packed(self, F):
        """Build every cached weight image (split input rows, packed recurrent rows, column sums) on the CURRENT stream
        with one 4-frame dry run of the encoder.  `separate` calls this before it forks its stream groups: the groups
        share the cache, and an entry built on group 0's stream would be read by the other groups' kernels with no
        ordering between the streams."""
        if self._packed_ready or self._tape is not None:
            return
        self._ensure_variables()
        T0 = max(4, int(getattr(self.encoder, 'TIME_ALIGN', 1) or 1))
        pending, self._stagger_pending = self._stagger_pending, False
        self._last_split = None
        self.encoder(torch.zeros((1, T0, F), dtype=torch.float32, device=self.device))
        self._last_split = None
        self._stagger_pending = pending
        self._packed_ready = True

    PIPELINE_GROUP = 8      # utterances per stream group = one recurrent cluster's batch tile
    LSTM_HEADSTART_US = 0   # experiment: delay the pipelined product by this much so the recurrent clusters are placed first
    STAGGER_US = 0          # experiment: release the next group this long after the first product starts (0 = when it ends)
    PREFETCH_H2D = True     # pinned host input: the groups' slices are copied in order by one copy stream
    PIPELINE_MAX_GROUPS = 4

    def separate(self, wav, groups=None, out=None):
        """demo path (main.py:660-695): waveforms [B,N] f32 -> [B,C,64*T] f32.
        STFT (app/utils.py:117-122), infer graph, iSTFT per source (app/utils.py:53-75).

        Utterances are independent, and the recurrence is latency-bound on a fraction of the SMs, so
        the batch is cut into groups that run front end / encoder / estimator / separator / back end on
        their own CUDA streams: one group's dense layers fill the SMs another group's recurrence leaves
        idle.  `wav` and `out` may be PINNED HOST tensors: each group then copies its own slice in and
        out on its stream, so the PCIe transfers of one group overlap the compute of the others."""
        B, n = wav.shape
        align = int(getattr(self.encoder, 'TIME_ALIGN', 1) or 1)
        if align > 1 and K.num_frames(n) % align:
            # conv-bilstm-v1 needs T % 4 == 0 (the reference pads its batches to LENGTH_ALIGN, main.py:667-668): append
            # silence up to the next aligned frame count and trim the result back to the 64*T samples of the input
            if out is not None or not wav.is_cuda:
                raise ValueError('the encoder needs a frame count that is a multiple of %d (got %d); pass a device tensor '
                                 'without `out` to have it padded' % (align, K.num_frames(n)))
            T_in = K.num_frames(n)
            T_al = T_in + (-T_in) % align
            padded = torch.nn.functional.pad(wav, (0, K.FFT_STRIDE * (T_al - 1) - n))
            return self.separate(padded, groups)[..., :K.FFT_STRIDE * T_in].contiguous()
        if groups is None:
            groups = max(1, min(self.PIPELINE_MAX_GROUPS, B // self.PIPELINE_GROUP))
            geo = getattr(self.encoder, '_geometry', None)
            if geo is not None and geo()[1] > K.TC_LSTM_MAX_H and not (
                    self.RECURRENT_FP16 and K.DEFAULT_BACKEND == 1 and geo()[1] <= K.TC_WIDE_MAX_H):
                # H too large for the tcgen05 kernels (or they are switched off): the fp32 cooperative kernel needs all
                # its CTAs resident, so more than two groups only queue behind each other (lstm-orig, H = 600: measured
                # 45.8 ms with 4, 31.6 with 2).  The wide tcgen05 kernel takes 19 SMs per group of 8: four groups fit.
                groups = min(groups, 2)
        Cn, T = hparams.MAX_N_SIGNAL, K.num_frames(n)
        if out is None:
            out = torch.empty((B, Cn, K.FFT_STRIDE * T), dtype=torch.float32, device=self.device)

        def run(lo, hi, w=None):
            K.stamp('g%d start' % lo)
            self._flag_pool = [K.pipeline_flags(self.device, self.FLAG_SETS), 0] if self.PIPELINE_INPUT_GEMM else None
            if w is None:
                w = wav[lo:hi]
                if not w.is_cuda:
                    w = w.to(self.device, non_blocking=True)
            mix, logmag = K.stft(w, want_logmag=True)
            K.stamp('g%d stft' % lo)
            fused = (self.USE_FUSED_K4 and getattr(self.separator, 'SUPPORTS_WAV', False) and Cn <= K.FUSED_K4_MAX_C
                     and hparams.EMBED_SIZE <= K.FUSED_K4_MAX_E and hparams.EMBED_SIZE % 4 == 0)
            if fused:
                wavs = self.infer(mix, logmag=logmag, want='wav', wav_out=out[lo:hi] if out.is_cuda else None)
                K.stamp('g%d mask' % lo)
                if not out.is_cuda:
                    out[lo:hi].copy_(wavs, non_blocking=True)
            else:
                sep = self.infer(mix, logmag=logmag)
                K.stamp('g%d mask' % lo)
                if out.is_cuda:
                    K.istft(sep, out=out[lo:hi])
                else:
                    out[lo:hi].copy_(K.istft(sep), non_blocking=True)
            K.stamp('g%d end' % lo)
            self._flag_pool = None

        if groups <= 1:
            run(0, B)
            return out
        main = torch.cuda.current_stream()
        self._prepare_packed(hparams.FEATURE_SIZE)      # shared weight images: built before the fork, on one stream
        fork = main.record_event()
        streams = self._side_streams(groups)
        staged = [None] * groups
        if not wav.is_cuda and self.PREFETCH_H2D:
            # host input: ONE copy stream moves the groups' slices in order, each at the full link rate, ahead of the
            # staggered starts (a group that copies its own slice when it starts delays every later group by the copy
            # time: measured +104 us on the last group's start; concurrent copies from four streams share the link and
            # delay the FIRST group instead)
            cp = self.side_stream('h2d')
            cp.wait_event(fork)
            with torch.cuda.stream(cp):
                for g in range(groups):
                    lo, hi = shard.shard_bounds(B, g, groups)
                    w = wav[lo:hi].to(self.device, non_blocking=True)
                    staged[g] = (w, cp.record_event())
        prev = None
        for g, st in enumerate(streams):
            lo, hi = shard.shard_bounds(B, g, groups)
            st.wait_event(fork)
            if staged[g] is not None:
                st.wait_event(staged[g][1])
            if prev is not None:
                # stagger: group g starts once group g-1 has queued its first dense layer, so the groups
                # do not march in lockstep (all dense, then all recurrent) but interleave the two phases
                st.wait_event(prev)
            with torch.cuda.stream(st):
                if prev is not None and self.STAGGER_US > 0:
                    torch.cuda._sleep(int(self.STAGGER_US * 1840))
                self._stagger_pending, self._stagger_event = True, None
                run(lo, hi, staged[g][0] if staged[g] is not None else None)
                prev = self._stagger_event
                self._stagger_pending = False
        for st in streams:
            main.wait_stream(st)
        return out

    def separate_graphed(self, wav, groups=None, out=None):
        """`separate` replayed from a CUDA graph captured once per (buffers, shape): ~400 kernel launches
        and their host-side argument checks collapse into one graph launch.  Device input: copied into a
        static buffer, the returned static output is overwritten by the next call.  Pinned host `wav` / `out`:
        the graph reads and writes those very buffers (their addresses are part of the key)."""
        host_io = not wav.is_cuda
        key = (tuple(wav.shape), groups, wav.data_ptr() if host_io else None, out.data_ptr() if out is not None else None)
        entry = self._graphs.get(key)
        if entry is None:
            static_in = wav if host_io else torch.empty_like(wav).copy_(wav)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):                     # warm up: variables, function attributes, allocator
                    self.separate(static_in, groups, out)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self.separate(static_in, groups, out)
            entry = (graph, static_in, static_out)
            self._graphs[key] = entry
        graph, static_in, static_out = entry
        if not host_io and wav.data_ptr() != static_in.data_ptr():
            static_in.copy_(wav, non_blocking=True)
        graph.replay()
        return static_out

    OVERLAP_WEIGHT_GRADS = True        # training: dW / db products on a side stream under the next layer's BPTT

    def side_stream(self, key):
        st = self._twins.get(key)
        if st is None:
            st = self._twins[key] = torch.cuda.Stream(device=self.device)
        return st

    LSTM_PRIORITY_STREAM = True        # measured: 2.862 -> 2.838 ms per step (tools/ab_groups.py)

    # Stream priorities of the grouped inference step: every recurrence launches from the highest level (its clusters
    # need whole SMs and carry the critical path); among the groups' own streams the EARLIER group outranks the later
    # one, so that at the end of the step -- four projections / mask kernels contending for the SMs the last
    # recurrences have not freed yet -- the groups finish in start order, a copy time apart, instead of in a convoy whose
    # device-to-host copies then queue on the link.
    GROUP_PRIORITIES = True

    def _priority_twin(self, stream):
        tw = self._twins.get(stream.cuda_stream)
        if tw is None:
            tw = self._twins[stream.cuda_stream] = torch.cuda.Stream(device=self.device,
                                                                     priority=-3 if self.GROUP_PRIORITIES else -1)
        return tw

    def _side_streams(self, n):
        pool = self._streams
        while len(pool) < n:
            g = len(pool)
            # torch's stream pools expose four levels (0 .. -3): the recurrences take -3, the groups -2, -1, 0, 0
            prio = min(0, -2 + g) if self.GROUP_PRIORITIES else 0
            pool.append(torch.cuda.Stream(device=self.device, priority=prio))
        return pool[:n]

    def separate_host(self, wav_host, out_host=None, graphed=True):
        """The call a user makes: pinned host waveforms in, (pinned) host waveforms out; returns after
        the results are in `out_host` only once the caller synchronises the current stream."""
        if out_host is None:
            n = wav_host.shape[1]
            out_host = torch.empty((wav_host.shape[0], hparams.MAX_N_SIGNAL, K.FFT_STRIDE * K.num_frames(n)),
                                   dtype=torch.float32).pin_memory()
        if graphed and wav_host.is_pinned() and out_host.is_pinned():
            return self.separate_graphed(wav_host, out=out_host)
        return self.separate(wav_host, out=out_host)
