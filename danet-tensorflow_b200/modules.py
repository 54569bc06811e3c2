"""
Encoder / Estimator / Separator plugins -- the reference's app/modules.py surface
(`cls(model, name)`, called like functions, registered with `@hparams.register_*`) on
CUDA tensors, every computation a call into libdanet_sm100.so (`kernels.py`).

Registered names: encoders `toy`, `lstm-orig`, `bilstm-orig`, `conv-bilstm-v1`; estimators `truth`,
`truth-threshold`, `truth-weighted`, `anchor`, `kmeans` (new); separators
`dot-sigmoid-orig`, `dot-softmax-orig`.
"""
from math import sqrt

import numpy as np

from . import kernels as K
from .hparams import hparams


class ModelModule(object):
    """app/modules.py:11-25"""
    def __init__(self, model, name):
        if hparams.DEBUG:
            self.debug_fetches = {}
        self.name = name
        self.model = model

    def __call__(self, s_dropout_keep=1.):
        raise NotImplementedError()


class Encoder(ModelModule):
    """app/modules.py:28-50: log-magnitude [B,T,F] -> embedding [B,T,F,E]"""
    def __call__(self, s_mixture, s_dropout_keep=1.):
        raise NotImplementedError()


class Estimator(ModelModule):
    """app/modules.py:53-70: embedding (+ truth) -> attractors [B,C,E]"""
    USE_TRUTH = True

    def __call__(self, s_embed, **kwargs):
        raise NotImplementedError()


class Separator(ModelModule):
    """app/modules.py:73-93: (|mix| [B,T,F], attractors [B,C,E], embed_flat [B,TF,E]) -> [B,C,T,F]"""
    def __call__(self, s_mixed_signals_pwr, s_attractors, s_embed_flat):
        raise NotImplementedError()


def lstm_bias_init(hdim):
    """app/modules.py:217-221 (and :157-161): [cand 0 | input 1.5 | forget -1 | output 1]"""
    b = np.zeros([hdim * 4], dtype=np.float32)
    b[hdim * 1:hdim * 2] = 1.5
    b[hdim * 2:hdim * 3] = -1.
    b[hdim * 3:hdim * 4] = 1.
    return b


def _uniform(lo, hi):
    return lambda rs, shape: rs.uniform(lo, hi, size=shape)


def _check_dropout(keep):
    # the reference never feeds its dropout placeholder into the encoder
    if keep != 1.:
        raise NotImplementedError('dropout is dead code in the reference (keep_prob is always 1)')


def _lyr_bilstm(name, model, s_input, hdim, w_init, b_init, s_dropout_keep=1.):
    """app/modules.py:120-137: forward scan, scan of the time-reversed input un-reversed, concat.
    Both directions run inside ONE recurrent kernel launch."""
    _check_dropout(s_dropout_keep)
    return model.lyr_bilstm(name, s_input, hdim, w_init=w_init, b_init=b_init)


def _glorot_uniform(rs, shape):
    lim = sqrt(6. / (shape[0] + shape[1]))          # TF1's default initialiser of tf.get_variable
    return rs.uniform(-lim, lim, size=shape)


def _zeros(rs, shape):
    return np.zeros(shape)


@hparams.register_encoder('toy')
class ToyEncoder(Encoder):
    """app/modules.py:96-116: 129 -> 2*FFT_SIZE -> F*E MLP with a leaky ReLU (the reference's default)"""
    def __call__(self, s_signals, s_dropout_keep=1.):
        model = self.model
        B, T, F = s_signals.shape
        E, hid = hparams.EMBED_SIZE, hparams.FFT_SIZE * 2
        W0 = model.get_variable('%s/linear0/W' % self.name, [F, hid], _glorot_uniform)
        B0 = model.get_variable('%s/linear0/B' % self.name, [hid], _zeros)
        W1 = model.get_variable('%s/linear1/W' % self.name, [hid, F * E], _glorot_uniform)
        B1 = model.get_variable('%s/linear1/B' % self.name, [F * E], _zeros)
        x2 = s_signals.reshape(B * T, F)
        mid = K.leaky_relu(K.linear(x2, W0, B0), hparams.RELU_LEAKAGE)
        if model._tape is not None:
            model._tape.append(dict(x=x2, mid=mid))
        return K.linear(mid, W1, B1).view(B, T, F, E)

    def backward(self, d_embed2, tape):
        """d_embed2 [B*T, F*E] -> fills model.grads (what tf.gradients derives for the two dense layers)"""
        g, P, n = self.model.grads, self.model.params, self.name
        rec = tape[0]
        K.gemm(rec['mid'], d_embed2, trans_a=True, out=g[n + '/linear1/W'])
        K.colsum(d_embed2, out=g[n + '/linear1/B'])
        dmid = K.leaky_relu_bwd(rec['mid'], K.gemm(d_embed2, P[n + '/linear1/W'], trans_b=True), hparams.RELU_LEAKAGE)
        K.gemm(rec['x'], dmid, trans_a=True, out=g[n + '/linear0/W'])
        K.colsum(dmid, out=g[n + '/linear0/B'])


def _recurrent_layer_backward(model, rec, dx, bidir, need_dx, main, side):
    """One (Bi)LSTM layer of the tape backwards: the BPTT kernel on the in-place gate gradients, then dX (critical path,
    returned) and, on `side`, the stacked [I+H, 4H] weight gradient and the bias gradient of every direction -- followed by
    the layer's bucket of the gradient exchange.  `dx` [B,T,n_dir*H] is the gradient of the layer's output."""
    g, P = model.grads, model.params
    H, x, I = rec['hdim'], rec['x'], rec['x'].shape[-1]
    B, T = x.shape[0], x.shape[1]
    names = [rec['name'] + '_fwd', rec['name'] + '_bwd'] if bidir else [rec['name']]
    Ws = [P[n + '/LSTM/linear/W'] for n in names]
    # wide layers (lstm-orig): the tcgen05 kernel with fp16 weights in the product where the forward already carried its
    # state as fp16 (Model.train_recurrent_fp16), else the exact fp32 kernel
    be = 2 if (K.DEFAULT_BACKEND == 1 and K.TC_LSTM_MAX_H < H <= K.TC_WIDE_MAX_H and model.train_recurrent_fp16()) else None
    if side is not main:
        # the clusters of the backward recurrence need whole SMs: queue them ahead of the side stream's tiles
        hp_stream = model._priority_twin(main)
        hp_stream.wait_stream(main)
        with K.torch.cuda.stream(hp_stream):
            da = K.lstm_seq_bwd(dx, rec['gates'], rec['cell'], Ws, I, T, B, H, backend=be)      # [n_dir,T,B,4H]
        main.wait_stream(hp_stream)
    else:
        da = K.lstm_seq_bwd(dx, rec['gates'], rec['cell'], Ws, I, T, B, H, backend=be)
    x2 = x.reshape(B * T, I)
    out2 = rec['out'].view(B * T, -1)
    dx_prev = None
    if need_dx:                                                                          # critical path first
        dx_prev = K.torch.empty((B * T, I), dtype=K.torch.float32, device=x.device)
        for d, n in enumerate(names):
            K.gemm(da[d].view(T * B, 4 * H), Ws[d][:I], trans_b=True, out_perm_T=B, out=dx_prev, accumulate=d > 0)
        dx_prev = dx_prev.view(B, T, I)
    side.wait_stream(main)                     # da is final (and, harmlessly, dX has been queued)
    with K.torch.cuda.stream(side):
        for d, n in enumerate(names):
            da_d = da[d].view(T * B, 4 * H)
            dW = g[n + '/LSTM/linear/W']
            # ONE product for the stacked [I+H, 4H] gradient: A = [x ; h shifted by one step] paired
            # time-major with da (sum_t X[b,t]^T da[t,b], sum_t h[b,t-+1]^T da[t,b]); each operand split once
            a2 = K.split_operand_paired(x2, T, 0, rows_total=I + H)
            K.split_operand_paired(out2[:, d * H:(d + 1) * H], T, -1 if d == 0 else 1, out=a2, row0=I,
                                   rows_total=I + H)
            b2 = K.split_operand(da_d, True)
            K.gemm_split(a2, b2, I + H, 4 * H, T * B, out=dW)
            K.colsum(da_d, out=g[n + '/LSTM/linear/B'])
        # this layer's gradients are queued: exchange them under the next layer's backward recurrence
        model.grads_ready([n + sfx for n in names for sfx in ('/LSTM/linear/W', '/LSTM/linear/B')])
    return dx_prev


class _RecurrentEncoder(Encoder):
    N_LAYERS = 4
    HDIM = 300
    INIT_SCALE = .75
    BIDIR = True
    TRAIN_FP16_STATE_OK = True     # Model.TRAIN_RECURRENT_FP16: the training forward may carry h as fp16 (measured, model.py)

    def _geometry(self):
        n_layers = getattr(hparams, 'ENCODER_LAYERS', None) or self.N_LAYERS
        hdim = getattr(hparams, 'ENCODER_HDIM', None) or self.HDIM
        return n_layers, hdim

    def __call__(self, s_signals, s_dropout_keep=1.):
        _check_dropout(s_dropout_keep)
        model = self.model
        B, T, F = s_signals.shape
        n_layers, hdim = self._geometry()
        x = K.center(s_signals)                                    # modules.py:209-210 / :150-151
        r = self.INIT_SCALE / sqrt(hdim)
        w_init = _uniform(-r, r)
        b_init = lambda rs, shape: lstm_bias_init(hdim)
        for l in range(n_layers):
            if self.BIDIR:
                # a layer that feeds another recurrent layer may hand its output over with time-major rows (model.py)
                model._emit_time_major = l + 1 < n_layers
                try:
                    x = _lyr_bilstm('%s/lstm%d' % (self.name, l), model, x, hdim, w_init, b_init, s_dropout_keep)
                finally:
                    model._emit_time_major = False
            else:
                x = model.lyr_lstm('%s/lstm%d' % (self.name, l), x, hdim, w_init=w_init, b_init=b_init)
        odim = x.shape[-1]
        E = hparams.EMBED_SIZE
        W = model.get_variable('%s/output/W' % self.name, [odim, F * E], _uniform(-1.85, 1.85))
        v = model.centered_projection('%s/output/W' % self.name, x, W)    # centring folded into the product
        if v is None:
            x = K.center(x)                                        # modules.py:244-245 / :181-182
            if model._tape is not None:
                model._tape.append(dict(centered=x))
            v = model.dense('%s/output/W' % self.name, x.view(B * T, odim), W)   # modules.py:249-255, no bias
        s_out = v.view(B, T, F, E)
        if hparams.DEBUG:
            self.debug_fetches['embed'] = s_out
        return s_out


    def backward(self, d_embed2, tape):
        """d_embed2 [B*T, F*E] -> fills model.grads: output projection, centring, then per layer the BPTT kernel
        and the dWx / dWh / db / dX products on its in-place gate gradients.

        Only  BPTT(l) -> dX(l) -> BPTT(l-1)  is a dependency chain; the weight and bias gradients of layer l are leaves.
        They run on a side stream under the next layer's BPTT, whose clusters occupy 80 of the 148 SMs and wait on
        DSMEM latency most of the time (measured: 13.1 -> see DESIGN.md section 7)."""
        model = self.model
        g, P = model.grads, model.params
        xc = tape[-1]['centered']
        B, T, odim = xc.shape
        main = K.torch.cuda.current_stream()
        # one side stream per training slice (Model.train_forward_backward in stream groups)
        side = model.side_stream(('grad', model._train_group)) if model.OVERLAP_WEIGHT_GRADS else main
        dx = K.gemm(d_embed2, P[self.name + '/output/W'], trans_b=True)                          # dX = dY W^T
        side.wait_stream(main)
        with K.torch.cuda.stream(side):
            K.gemm(xc.view(B * T, odim), d_embed2, trans_a=True, out=g[self.name + '/output/W'])  # dW = X^T dY
            # first bucket of the gradient exchange: the projection and (written before this backward) the anchors
            model.grads_ready([self.name + '/output/W'] + [k for k in g if k.endswith('/anchors')])
        dx = K.center(dx.view(B, T, odim))             # the gradient of x - mean(x) is the same centring
        for l in range(len(tape) - 2, -1, -1):
            dx = _recurrent_layer_backward(model, tape[l], dx, self.BIDIR, need_dx=l > 0, main=main, side=side)
        main.wait_stream(side)


@hparams.register_encoder('lstm-orig')
class LstmEncoder(_RecurrentEncoder):
    """app/modules.py:140-196: 4 x 600 unidirectional"""
    HDIM = 600
    INIT_SCALE = 1.15
    BIDIR = False


@hparams.register_encoder('bilstm-orig')
class BiLstmEncoder(_RecurrentEncoder):
    """app/modules.py:199-260: 4 x (300 + 300)"""


def _glorot_uniform_conv(rs, shape):
    # tf.layers.conv2d default kernel initialiser on [k, k, cin, cout]: the fans include the receptive field
    rf = shape[0] * shape[1]
    lim = sqrt(6. / (rf * shape[2] + rf * shape[3]))
    return rs.uniform(-lim, lim, size=shape)


@hparams.register_encoder('conv-bilstm-v1')
class ConvBiLstmEncoder(Encoder):
    """app/modules.py:263-379, the reference's experimental CNN-LSTM hybrid: two conv + max-pool stages
    (1 -> 8 -> 16 channels, then 16 -> 32 -> 16), a 2-layer BiLSTM (2 x FFT_SIZE hidden units) over the T/4 pooled
    frames with a residual connection, two 3x3 convs whose 64 channels are un-pooled by depth-to-space, two 5x5 convs,
    and a bias-free dense layer FFT_SIZE -> F*E.  T must be a multiple of 4 (the reference pads, main.py:667-671).
    Variables are created under tf.layers' names (conv2d, conv2d_1, ..., dense) in the reference's order.
    Trains too: `backward` is the hand-derived adjoint of the whole stack, pinned by the reference's own tf.gradients
    values (tests/golden/model_convbilstm_anchor_softmax.npz)."""
    TIME_ALIGN = 4             # two 2x2 max-pools over time: Model.separate pads waveforms to a multiple of 4 frames

    def _conv(self, idx, x, k, cout, init=None):
        model = self.model
        nm = '%s/conv2d%s' % (self.name, '' if idx == 0 else '_%d' % idx)
        w = model.get_variable(nm + '/kernel', [k, k, x.shape[1], cout], init or _glorot_uniform_conv)
        b = model.get_variable(nm + '/bias', [cout], _zeros)
        y = K.conv2d(x, w, b, leak=hparams.RELU_LEAKAGE)
        if model._tape is not None:
            self._saved['conv%d' % idx] = (nm, x, y)
        return y

    def __call__(self, s_signals, s_dropout_keep=1.):
        _check_dropout(s_dropout_keep)
        model = self.model
        B, T, F = s_signals.shape
        nfft, E = hparams.FFT_SIZE, hparams.EMBED_SIZE
        if T % 4 != 0:
            raise ValueError('conv-bilstm-v1 needs a frame count that is a multiple of 4 (got %d)' % T)
        if F != nfft // 2 + 1:
            raise ValueError('conv-bilstm-v1: FEATURE_SIZE %d does not match FFT_SIZE %d' % (F, nfft))
        r = 2. / sqrt(nfft)
        w_init = _uniform(-r, r)
        self._saved = {}

        def b_init(rs, shape):                                   # :280-285: [0 | input 1 | forget -1 | output 1]
            b = np.zeros(shape)
            b[nfft:2 * nfft], b[2 * nfft:3 * nfft], b[3 * nfft:] = 1., -1., 1.
            return b

        x = s_signals.reshape(B, 1, T, F)
        m0 = self._conv(0, x, 5, 8)                                               # :289-293
        p0_in = self._conv(1, m0, 5, 16)
        m0 = K.maxpool2x2(p0_in)                                                  # :294-300  [B,16,T/2,64]
        m1 = self._conv(2, m0, 3, 32)
        p1_in = self._conv(3, m1, 3, 16)
        m1 = K.maxpool2x2(p1_in)                                                  # :302-313  [B,16,T/4,32]
        T4, F8 = m1.shape[2], m1.shape[3]
        m1 = K.center(m1.view(B, 16 * T4, F8)).view(B, 16, T4, F8)                # :315
        m2 = m1.permute(0, 2, 1, 3).reshape(B, T4, nfft * 2)                      # :318-320
        m2 = _lyr_bilstm('%s/lstm0' % self.name, model, m2, nfft, w_init, b_init)
        m3 = _lyr_bilstm('%s/lstm1' % self.name, model, m2, nfft, w_init, b_init)
        m3 = m3.reshape(B, T4, 16, F8).permute(0, 2, 1, 3).contiguous()           # :331-333
        m3 = K.add(m3, m1)                                                        # :335
        m3 = K.center(m3.view(B, 16 * T4, F8)).view(B, 16, T4, F8)                # :336
        m4 = self._conv(4, m3, 3, 32, _uniform(-.3, .3))
        m4 = self._conv(5, m4, 3, 64, _uniform(-.3, .3))                          # :342-353  [B,64,T/4,32]
        m4 = m4.view(B, 16, 2, 2, T4, F8).permute(0, 1, 4, 2, 5, 3).reshape(B, 16, 2 * T4, 2 * F8)   # :354-357
        m5 = self._conv(6, m4, 5, 16)
        m5 = self._conv(7, m5, 5, 8)                                              # :359-369  [B,8,T/2,64]
        m5 = m5.permute(0, 2, 1, 3).reshape(B * T, nfft)                          # :371-373
        W = model.get_variable('%s/dense/kernel' % self.name, [nfft, F * E], _glorot_uniform)
        s_out = model.dense('%s/dense/kernel' % self.name, m5, W).view(B, T, F, E)                  # :375-379
        if model._tape is not None:
            self._saved.update(pool0_in=p0_in, pool1_in=p1_in, m5=m5, geom=(B, T, T4, F8))
            model._tape.append(dict(conv=self._saved))
        if hparams.DEBUG:
            self.debug_fetches.update(conv_act=m1, lstm_act=m3, mid4=m4)
        return s_out

    def backward(self, d_embed2, tape):
        """d_embed2 [B*T, F*E] -> fills model.grads with what tf.gradients derives for this encoder (main.py:357-358):
        dense layer, the four back-end convolutions through the depth-to-space shuffle, the residual sum and both
        centrings, the two BiLSTM layers (BPTT kernels, as bilstm-orig), the two max-pools and the four front-end
        convolutions.  The convolution gradients run on direct fp32 kernels (conv.cu)."""
        model = self.model
        g, P, name = model.grads, model.params, self.name
        sv = tape[-1]['conv']
        lstm = [rec for rec in tape if 'gates' in rec]
        B, T, T4, F8 = sv['geom']
        nfft = hparams.FFT_SIZE
        leak = hparams.RELU_LEAKAGE
        main = K.torch.cuda.current_stream()

        def conv_bwd(idx, dy, need_dx=True):
            nm, x, y = sv['conv%d' % idx]
            dx, dw, db = K.conv2d_bwd(x, P[nm + '/kernel'], y, dy.contiguous(), leak=leak, need_dx=need_dx)
            g[nm + '/kernel'].copy_(dw)
            g[nm + '/bias'].copy_(db)
            return dx

        Wd = P[name + '/dense/kernel']
        K.gemm(sv['m5'], d_embed2, trans_a=True, out=g[name + '/dense/kernel'])               # dW = m5^T dY
        d5 = K.gemm(d_embed2, Wd, trans_b=True)                                              # [B*T, nfft]
        d5 = d5.view(B, T // 2, 8, 2 * F8).permute(0, 2, 1, 3)                               # inverse of :371-373
        d4 = conv_bwd(6, conv_bwd(7, d5))                                                    # [B,16,T/2,64]
        d4 = d4.view(B, 16, T4, 2, F8, 2).permute(0, 1, 3, 5, 2, 4).reshape(B, 64, T4, F8)   # inverse of :354-357
        d3 = conv_bwd(4, conv_bwd(5, d4))                                                    # [B,16,T/4,32]
        d3 = K.center(d3.contiguous().view(B, 16 * T4, F8)).view(B, 16, T4, F8)              # :336 (self-adjoint)
        d_m1 = d3                                                                            # residual branch (:335)
        dx = d3.permute(0, 2, 1, 3).reshape(B, T4, 2 * nfft).contiguous()                    # inverse of :331-333
        for l in (1, 0):
            dx = _recurrent_layer_backward(model, lstm[l], dx, True, need_dx=True, main=main, side=main)
        d_m1 = K.add(d_m1.contiguous(), dx.view(B, T4, 16, F8).permute(0, 2, 1, 3).contiguous())   # inverse of :318-320
        d_m1 = K.center(d_m1.view(B, 16 * T4, F8)).view(B, 16, T4, F8)                       # :315
        d = K.maxpool2x2_bwd(sv['pool1_in'], d_m1)
        d = conv_bwd(2, conv_bwd(3, d))
        d = K.maxpool2x2_bwd(sv['pool0_in'], d)
        conv_bwd(0, conv_bwd(1, d), need_dx=False)
        # (the convolution / dense gradients are exchanged by GradientBuckets.finish(): whatever no bucket covered)


class _TruthEstimator(Estimator):
    USE_TRUTH = True
    MODE = 'truth'

    def __call__(self, s_embed, s_src_pwr=None, s_mix_pwr=None, s_embed_flat=None):
        if s_src_pwr is None:
            raise ValueError('estimator "%s" needs the true source magnitudes' % self.MODE)
        return K.attractor_truth(s_embed, s_src_pwr, s_mix_pwr, self.MODE)


@hparams.register_estimator('truth')
class AverageEstimator(_TruthEstimator):
    """app/modules.py:382-412"""
    MODE = 'truth'


@hparams.register_estimator('truth-threshold')
class ThreshouldedAverageEstimator(_TruthEstimator):
    """app/modules.py:415-450"""
    MODE = 'truth-threshold'


@hparams.register_estimator('truth-weighted')
class WeightedAverageEstimator(_TruthEstimator):
    """app/modules.py:453-487"""
    MODE = 'truth-weighted'


def _normal(rs, shape):
    return rs.standard_normal(size=shape)


@hparams.register_estimator('anchor')
class AnchoredEstimator(Estimator):
    """app/modules.py:490-545"""
    USE_TRUTH = False

    def anchors(self):
        return self.model.get_variable('%s/anchors' % self.name, [hparams.NUM_ANCHOR, hparams.EMBED_SIZE], _normal)

    def __call__(self, s_embed, s_src_pwr=None, s_mix_pwr=None, s_embed_flat=None):
        fused = self.model._fused_attrs
        if fused is not None:
            # the encoder's output projection already took this estimator's sums in its epilogue (Model.centered_projection)
            self.model._fused_attrs = None
            if type(self) is AnchoredEstimator and fused[0].data_ptr() == s_embed.data_ptr():
                return fused[1]
        if hparams.DEBUG:
            out, sets, sims, choice = K.attractor_anchor(s_embed, self.anchors(), hparams.MAX_N_SIGNAL, True)
            self.debug_fetches.update(asets=sets, anchors=self.anchors(), subset_choice=choice)
            return out
        return K.attractor_anchor(s_embed, self.anchors(), hparams.MAX_N_SIGNAL)


@hparams.register_estimator('kmeans')
class KMeansEstimator(AnchoredEstimator):
    """NEW plugin (the reference's README.md:216-217 lists k-means as unimplemented):
    Lloyd iterations seeded with the anchor estimator's attractors."""
    USE_TRUTH = False
    N_ITER = 5

    def __call__(self, s_embed, s_src_pwr=None, s_mix_pwr=None, s_embed_flat=None):
        init = K.attractor_anchor(s_embed, self.anchors(), hparams.MAX_N_SIGNAL)
        return K.attractor_kmeans(s_embed, init, getattr(hparams, 'KMEANS_ITERS', None) or self.N_ITER)


class _DotSeparator(Separator):
    KIND = None

    SUPPORTS_WAV = True      # want=('wav',): the fused back end

    def __call__(self, s_mixed_signals_pwr, s_attractors, s_embed_flat, s_mixed_signals=None, want=('sep_pwr',),
                 s_wav_out=None):
        """The reference signature returns magnitudes [B,C,T,F].  Extension: passing the complex
        mixture and `want` also yields the re-phased spectra (main.py:281-284) and the masks from the
        same pass."""
        if tuple(want) == ('wav',):
            # demo path (main.py:660-695): straight to waveforms, mask x mixture -> iSTFT in one kernel
            return dict(wav=K.mask_cmul_istft(s_embed_flat, s_attractors, s_mixed_signals, self.KIND, out=s_wav_out))
        out = K.mask_cmul(s_embed_flat, s_attractors, s_mixed_signals, self.KIND, want=want,
                          mix_pwr=s_mixed_signals_pwr)
        if hparams.DEBUG and out.get('masks') is not None:
            self.debug_fetches['masks'] = out['masks']
        if tuple(want) == ('sep_pwr',):
            return out['sep_pwr']
        return out


@hparams.register_separator('dot-sigmoid-orig')
class DotSeparatorSigmoid(_DotSeparator):
    """app/modules.py:548-574"""
    KIND = 'dot-sigmoid-orig'


@hparams.register_separator('dot-softmax-orig')
class DotSeparatorSoftmax(_DotSeparator):
    """app/modules.py:577-603"""
    KIND = 'dot-softmax-orig'
