"""Drop-in for mnist/model.py: class DCGAN with build_model / train / discriminator / generator /
gen_sampler / classifier / load_mnist, the same constructor arguments and flag semantics.

`build_model()` records two static programs (graph.py) instead of a TF graph:
  d_prog : G(z,y_gen) forward, D(real), D(fake), classifier(real)   -> gradients w.r.t. d_vars
  g_prog : G(z,y_gen) forward+backward, D(fake) (data gradients only), classifier(G)
           -> gradients w.r.t. g_vars and confusion_logits
and `train_iteration()` is the body of the reference's hot loop (mnist/model.py:335-372): one D step
then two G(+C) steps on the same z / labels, each step a captured CUDA graph.

Where the reference evaluates the discriminator 10 times with the 10 one-hot labels (unbiased real
branch :153-174, learned-C fake branch :183-204) the label only enters through the projection head
h6 = h4 + <h3, h5(y)>, so ONE trunk evaluation plus the fused channel loss gives identical values
(checked against the literal 10-call oracle in tests/).
"""
import math
import os
import time
from types import SimpleNamespace

import numpy as np
import torch

from . import _C, ops, scope as S
from .graph import Program, VariableStore, tf_adam_lr
from .nnops import (CastOp, ChannelLossOp, ConcatRowsOp, ConvOp, LogitLossOp, MeanHWOp, RecoverMSEOp, SigmoidCEOp, SoftmaxRowsOp, adam_step)
from .sampler import LabelNoiseSampler, class_dependent_confusion, one_coin_confusion

SPLIT_GRAPH = __import__('os').environ.get('RCGAN_DP_SPLIT_GRAPH', '0') == '1'   # A/B: blocking all-reduce between two graph halves

LOSS_MODES = {'hinge': (_C.HINGE_D_REAL, _C.HINGE_D_FAKE, _C.HINGE_G), 'ce': (_C.CE_D_REAL, _C.CE_D_FAKE, _C.CE_G)}


def default_flags(**kw):
    """Flag defaults of mnist/main.py:13-66 (hot-path subset; unknown names are ignored there too)."""
    f = SimpleNamespace(
        epoch=5, learning_rate=2e-4, beta1=0.5, train_size=np.inf, batch_size=100, dataset='mnist', z_dim=100,
        algorithm='biased', estimate_confuse=True, confuse_multiplier=10.0, perm_regularizer=True, perm_multiplier=10.0,
        alpha=1.0, confusion_class_depend=False, disc_type='vanilla', loss_fn='hinge', real_match=False, add_noise=False,
        noise_alpha=0.3, noise_start=30, noise_end=80, concat_y=False, concat_y_layers=[1], spectral_norm=True,
        max_norm=True)
    for k, v in kw.items():
        setattr(f, k, v)
    f.concat_y_layers = [int(x) for x in f.concat_y_layers]
    return f


def conv_out_size_same(size, stride):
    return int(math.ceil(float(size) / float(stride)))


def _group_of(name):
    """mnist/model.py:242-245 variable partition by substring; confusion_logits has its own optimizer."""
    if name == 'confusion_logits':
        return 'c'
    if 'd_' in name:
        return 'd'
    if 'g_' in name:
        return 'g'
    return None


class DCGAN(object):
    def __init__(self, sess=None, input_height=28, input_width=28, crop=False, batch_size=64, sample_num=64,
                 output_height=28, output_width=28, y_dim=10, z_dim=100, gf_dim=64, df_dim=64, gfc_dim=1024, dfc_dim=1024,
                 c_dim=1, dataset_name='mnist', checkpoint_dir=None, sample_dir=None, data_dir='./data', algorithm='biased',
                 estimate_confuse=False, perm_regularizer=False, alpha=1.0, disc_type='vanilla', add_noise=False,
                 noise_alpha=1.0, config=None, device='cuda', precision='bf16', seed=0, data=None, world_size=1, rank=0,
                 use_cuda_graph=True, pre_norm_fp32=False):
        """Arguments up to `config` are the reference's (mnist/model.py:19-26); `sess` is vestigial.
        precision: 'bf16' (bf16 activations, fp32 accumulate / parameters) or 'fp32' (parity mode).
        data: optional (X [N,28,28,1] float in [0,1], y [N] int) -- the reference reads the MNIST idx
        files here; datasets are out of scope (synthetic inputs), so data is passed in or fed per batch."""
        self.algorithm, self.estimate_confuse, self.perm_regularizer = algorithm, estimate_confuse, perm_regularizer
        self.alpha, self.disc_type, self.add_noise, self.noise_alpha = alpha, disc_type, add_noise, noise_alpha
        self.sess, self.crop = sess, crop
        self.batch_size, self.sample_num = batch_size, sample_num
        self.input_height, self.input_width = input_height, input_width
        self.output_height, self.output_width = output_height, output_width
        self.y_dim, self.z_dim = y_dim, z_dim
        self.gf_dim, self.df_dim, self.gfc_dim, self.dfc_dim, self.c_dim = gf_dim, df_dim, gfc_dim, dfc_dim, c_dim
        self.config = config if config is not None else default_flags(
            algorithm=algorithm, estimate_confuse=estimate_confuse, perm_regularizer=perm_regularizer, alpha=alpha,
            disc_type=disc_type, batch_size=batch_size)
        self.dataset_name, self.checkpoint_dir, self.data_dir = dataset_name, checkpoint_dir, data_dir
        self.device = torch.device(device)
        self.precision = precision
        self.act_dtype = {'bf16': _C.BF16, 'fp32': _C.F32}[precision]
        self.seed = seed
        self.world_size, self.rank = world_size, rank
        self.use_cuda_graph = use_cuda_graph
        # bf16 mode only: store the tensors that feed a batch norm in fp32 (halves the rounding that the norm's
        # backward amplifies; costs 2 extra bytes per element of traffic).  See DESIGN.md 'bf16 conditioning'.
        self.pre_norm_fp32 = pre_norm_fp32
        if not _C.load().rcgan_device_ok():
            raise _C.RcganError('rcgan_b200 needs an sm_100 device: ' + _C.last_error())

        self.d_bn1 = ops.batch_norm(name='d_bn1')
        self.d_bn2 = ops.batch_norm(name='d_bn2')
        if self.y_dim:
            self.d_bn3 = ops.batch_norm(name='d_bn3')
        self.g_bn0 = ops.batch_norm(name='g_bn0')
        self.g_bn1 = ops.batch_norm(name='g_bn1')
        self.g_bn2 = ops.batch_norm(name='g_bn2')

        if self.config.confusion_class_depend:
            self.confusion_matrix_actual = class_dependent_confusion(self.alpha)
        else:
            self.confusion_matrix_actual = one_coin_confusion(self.alpha, self.y_dim)
        if data is not None:
            (self.data_X, self.data_y_actual, self.data_y_real, self.data_y_gen, self.data_y_fake,
             self.data_y_real_weights) = self.load_mnist(data)
        self.grayscale = (self.c_dim == 1)
        self.build_model()

    # ------------------------------------------------------------------ data / sampler
    def load_mnist(self, data):
        """mnist/model.py:770-834 minus the idx-file IO: seed 547, shuffle X and y with the same stream,
        then the per-sample label-noise sampler -- on the device, bit-exact with numpy's legacy stream."""
        X, y = data
        smp = LabelNoiseSampler(self.device)
        out = smp.load_mnist_labels(np.asarray(y), self.confusion_matrix_actual, real_match=self.config.real_match,
                                    seed=547, shuffle=True)
        self.sampler_state = smp
        X = np.asarray(X)[out['perm']]
        eye = np.eye(self.y_dim)
        C_inv = np.linalg.inv(self.confusion_matrix_actual)
        return (X, eye[out['y']], eye[out['real']], eye[out['gen']], eye[out['fake']], C_inv[out['real']])

    # ------------------------------------------------------------------ networks
    def generator(self, z, y=None, train=True):
        """mnist/model.py:705-731 (train=False: gen_sampler :733-757)."""
        with S.variable_scope("generator"):
            s_h, s_w = self.output_height, self.output_width
            s_h2, s_h4 = int(s_h / 2), int(s_h / 4)
            s_w2, s_w4 = int(s_w / 2), int(s_w / 4)
            B = self.batch_size
            z = ops.concat_label(z, y)
            # (the concat([h, y]) / conv_cond_concat after g_bn0 and g_bn2 is written by the norm's apply pass itself)
            h0 = self.g_bn0(ops.linear(z, self.gfc_dim, 'g_h0_lin', pre_norm=self.pre_norm_fp32), train=train, fuse_act='relu',
                            concat_y=y)
            h1 = self.g_bn1(ops.linear(h0, self.gf_dim * 2 * s_h4 * s_w4, 'g_h1_lin', pre_norm=self.pre_norm_fp32), train=train,
                            fuse_act='relu')
            h1 = h1.view([B, s_h4, s_w4, self.gf_dim * 2])
            h1 = ops.conv_cond_concat(h1, y)
            h2 = self.g_bn2(ops.deconv2d(h1, [B, s_h2, s_w2, self.gf_dim * 2], name='g_h2', pre_norm=self.pre_norm_fp32), train=train,
                            fuse_act='relu', concat_y=y)
            return ops.deconv2d(h2, [B, s_h, s_w, self.c_dim], name='g_h3', fuse_act='sigmoid')

    def gen_sampler(self, z, y=None):
        return self.generator(z, y, train=False)

    def _d_trunk(self, image, y, groups=1):
        """Label-independent part of the projection discriminator (mnist/model.py:649-678) -> h3 [B, df_dim].
        groups=2: `image` is [real; fake] and every batch norm keeps separate statistics per half."""
        cfg = self.config
        cc = lambda l, t: ops.conv_cond_concat(t, y) if (cfg.concat_y and l in cfg.concat_y_layers) else t
        sn = cfg.spectral_norm
        h0 = ops.conv2d(cc(1, image), self.df_dim, spectral_norm=sn, name='d_h0_conv', fuse_act='lrelu')
        h1 = self.d_bn1(ops.conv2d(cc(2, h0), self.df_dim, spectral_norm=sn, name='d_h1_conv', pre_norm=self.pre_norm_fp32), fuse_act='lrelu',
                        track_moving=False, groups=groups)
        h2 = self.d_bn2(ops.conv2d(cc(3, h1), self.df_dim, spectral_norm=sn, name='d_h2_conv', pre_norm=self.pre_norm_fp32), fuse_act='lrelu',
                        track_moving=False, groups=groups)
        h3 = self.d_bn3(ops.conv2d(cc(4, h2), self.df_dim, spectral_norm=sn, name='d_h3_conv', pre_norm=self.pre_norm_fp32), fuse_act='lrelu',
                        track_moving=False, groups=groups)
        return MeanHWOp(h3).y

    def discriminator_pair(self, x_real, x_fake, y_real, y_fake, wgt_real, wgt_fake, modes, names):
        """The D step's two projection-discriminator calls (mnist/model.py:150-207: D(x, .) and D(G(z), .)) as ONE pass over
        [real; fake]: convs, linears and their weight gradients see 2B samples per launch, the batch norms keep the two
        calls' separate batch statistics (groups=2), and each half gets its own loss term / channel weights.
        Mathematically identical to two calls (every other op is per-sample).  Returns (logits_real, logits_fake)."""
        B = self.batch_size
        with S.variable_scope("discriminator"):
            x_all = ConcatRowsOp(x_real, x_fake).y
            y_all = ConcatRowsOp(y_real, y_fake).y if self.config.concat_y else None
            h3 = self._d_trunk(x_all, y_all, groups=2)
            h3f = CastOp(h3, _C.F32).y                       # fp32 head
            h4 = ops.linear(h3f, 1, 'd_h4_lin', max_norm=self.config.max_norm)
            V = ops.linear(self._eye(Program.current), self.df_dim, 'd_h5_y_lin', max_norm=self.config.max_norm)
            op_r = ChannelLossOp(h3f, h4, V, wgt_real, modes[0], names[0], rows=(0, B))
            op_f = ChannelLossOp(h3f, h4, V, wgt_fake, modes[1], names[1], rows=(B, B))
            return op_r.logits, op_f.logits

    def discriminator(self, image, y=None, reuse=False, wgt=None, mode=None, loss_name=None, coef=1.0):
        """mnist/model.py:644-703.  Records the discriminator on `image` and its GAN loss term.
        projection: y is ignored by the trunk (unless concat_y) and enters through `wgt` [B,10], the
        per-class weights of the channel loss (one-hot y for a plain D(x,y)); returns the [B,10] tensor of
        logits for every class.  vanilla: y [B,10] is concatenated as in the reference; returns [B,1] logits."""
        prog = Program.current
        B = self.batch_size
        with S.variable_scope("discriminator"):
            if self.disc_type == "projection":
                h3 = self._d_trunk(image, y)
                h3f = CastOp(h3, _C.F32).y                       # fp32 head
                h4 = ops.linear(h3f, 1, 'd_h4_lin', max_norm=self.config.max_norm)
                # d_h5_y_lin applied to the 10 one-hot labels at once: V[j,:] = Matrix[j,:] + bias
                V = ops.linear(self._eye(prog), self.df_dim, 'd_h5_y_lin', max_norm=self.config.max_norm)
                op = ChannelLossOp(h3f, h4, V, wgt if wgt is not None else y, mode, loss_name, coef)
                return op.logits
            yb = y
            x = ops.conv_cond_concat(image, yb)
            h0 = ops.conv2d(x, self.c_dim + self.y_dim, name='d_h0_conv', fuse_act='lrelu')
            h0 = ops.conv_cond_concat(h0, yb)
            h1 = self.d_bn1(ops.conv2d(h0, self.df_dim + self.y_dim, name='d_h1_conv', pre_norm=self.pre_norm_fp32), fuse_act='lrelu',
                            track_moving=False)
            h1 = h1.view([B, h1.rows * h1.c // B])
            h1 = CastOp(h1, _C.F32).y
            h1 = ops.concat_label(h1, y)
            h3 = self.d_bn2(ops.linear(h1, self.dfc_dim, 'd_h3_lin'), fuse_act='lrelu', track_moving=False)
            h3 = ops.concat_label(h3, y)
            h4 = ops.linear(h3, 1, 'd_h4_lin')
            LogitLossOp(h4, mode, loss_name, coef)
            return h4

    def classifier(self, x, reuse=False):
        """mnist/model.py:759-768; x must be fp32 [B,h,w,c]."""
        with S.variable_scope("classifier"):
            return ops.linear(x.view([self.batch_size, x.rows * x.c // self.batch_size]), self.y_dim, 'd_classifier_h1')

    def _eye(self, prog):
        if not hasattr(prog, '_eye'):
            prog._eye = prog.new((self.y_dim, self.y_dim), _C.F32)
            prog._eye.data.copy_(torch.eye(self.y_dim, device=self.device).reshape(-1))
        return prog._eye

    def _confusion(self, prog):
        """mnist/model.py:102-108."""
        if self.estimate_confuse:
            logits = S.get_variable('confusion_logits', [self.y_dim, self.y_dim], self._glorot)
            return SoftmaxRowsOp(logits).y
        C = prog.new((self.y_dim, self.y_dim), _C.F32)
        C.data.copy_(torch.as_tensor(self.confusion_matrix_actual, dtype=torch.float32).reshape(-1))
        return C

    @staticmethod
    def _glorot(shape):
        lim = math.sqrt(6.0 / (shape[0] + shape[1]))
        return (torch.rand(tuple(shape), generator=S.init_generator(), dtype=torch.float32) * 2 - 1) * lim

    # ------------------------------------------------------------------ graph
    def build_model(self):
        cfg = self.config
        B, dev = self.batch_size, self.device
        if cfg.concat_y and self.disc_type == 'projection' and (self.algorithm == 'unbiased' or (
                self.algorithm in ('rcgan', 'ambient') and self.estimate_confuse)):
            raise NotImplementedError('label-concat trunk with a 10-label expectation is not used by any reference '
                                      'run script and is not built')
        if self.algorithm not in ('biased', 'rcgan', 'ambient', 'unbiased'):
            raise ValueError('Unknown algorithm: {}'.format(self.algorithm))
        if cfg.loss_fn not in LOSS_MODES:
            raise ValueError('Unknown self.config.loss_fn: {}!'.format(cfg.loss_fn))
        m_real, m_fake, m_gen = LOSS_MODES[cfg.loss_fn]
        image_dims = [self.output_height, self.output_width, self.c_dim] if self.crop else \
            [self.input_height, self.input_width, self.c_dim]
        self.store = VariableStore(dev, _group_of)
        S.set_store(self.store, self.seed)
        proj = self.disc_type == 'projection'
        learned = self.algorithm in ('rcgan', 'ambient') and self.estimate_confuse

        def fake_weights(prog, y_gen, y_fake):
            """which per-class weights the fake branch uses (mnist/model.py:176-207)"""
            if self.algorithm in ('rcgan', 'ambient'):
                if not self.estimate_confuse:
                    return y_fake
                C = self._confusion(prog)
                return ConvOp(y_gen, C, None).y             # tensordot(y_gen, confusion_matrix)
            return y_gen

        # ---------------- D step program
        self.d_prog = dp_ = Program('d_step', dev, self.act_dtype)
        with dp_:
            x32 = dp_.input('inputs', [B] + image_dims)
            z = dp_.input('z', [B, self.z_dim])
            y_real = dp_.input('y_real', [B, self.y_dim])
            y_gen = dp_.input('y_gen', [B, self.y_dim])
            y_fake = dp_.input('y_fake', [B, self.y_dim])
            y_rw = dp_.input('y_real_weights', [B, self.y_dim])
            x = CastOp(x32, self.act_dtype).y
            zc = CastOp(z, self.act_dtype).y
            G = self.generator(zc, y_gen)
            if learned and not proj:
                raise NotImplementedError('estimate_confuse needs the projection discriminator (run_rcganu.sh)')
            if self.algorithm == 'unbiased' and not proj:
                raise NotImplementedError('unbiased needs the projection discriminator (run_unbiased.sh)')
            wf = fake_weights(dp_, y_gen, y_fake)
            w_real = y_rw if self.algorithm == 'unbiased' else y_real
            if proj and not (learned and cfg.concat_y):
                self.D_logits, self.D_logits_ = self.discriminator_pair(x, G, y_real, wf, w_real, wf, (m_real, m_fake),
                                                                        ('d_loss_real', 'd_loss_fake'))
            else:
                if self.algorithm == 'unbiased':
                    self.D_logits = self.discriminator(x, y_real, wgt=y_rw, mode=m_real, loss_name='d_loss_real')
                else:
                    self.D_logits = self.discriminator(x, y_real, mode=m_real, loss_name='d_loss_real')
                self.D_logits_ = self.discriminator(G, None if learned else wf, reuse=True, wgt=wf, mode=m_fake,
                                                    loss_name='d_loss_fake')
            if self.perm_regularizer:
                self.classifier_logits = self.classifier(x32)
                SigmoidCEOp(self.classifier_logits, y_real, 'class_loss_real', 1.0)
            self.G = G
        # ---------------- G step program
        self.g_prog = gp_ = Program('g_step', dev, self.act_dtype)
        with gp_:
            z = gp_.input('z', [B, self.z_dim])
            y_gen = gp_.input('y_gen', [B, self.y_dim])
            y_fake = gp_.input('y_fake', [B, self.y_dim])
            zc = CastOp(z, self.act_dtype).y
            Gg = self.generator(zc, y_gen)
            wf = fake_weights(gp_, y_gen, y_fake)
            self.discriminator(Gg, None if learned else wf, reuse=True, wgt=wf, mode=m_gen, loss_name='g_loss')
            if self.perm_regularizer:
                G32 = CastOp(Gg, _C.F32).y
                SigmoidCEOp(self.classifier(G32, reuse=True), y_gen, 'class_loss_fake', cfg.perm_multiplier)
            self.G_gstep = Gg
        # ---------------- sampler program (gen_sampler: BN in inference mode, mnist/model.py:733-757)
        self.s_prog = sp_ = Program('sampler', dev, self.act_dtype)
        with sp_:
            z = sp_.input('z', [B, self.z_dim])
            y_gen = sp_.input('y_gen', [B, self.y_dim])
            self.sampler = CastOp(self.gen_sampler(CastOp(z, self.act_dtype).y, y_gen), _C.F32).y

        self.store.finalize()
        V = self.store.vars
        self.d_vars = [v for n, v in V.items() if v.trainable and 'd_' in n]
        self.g_vars = [v for n, v in V.items() if v.trainable and 'g_' in n]
        self.c_vars = [V['confusion_logits']] if 'confusion_logits' in V else []
        self.d_prog.finalize(self.d_vars)
        self.g_prog.finalize(self.g_vars + self.c_vars)
        self.s_prog.finalize([])
        G_ = self.store.groups
        self.groups = {'d': G_['d'], 'g': G_['g']}
        if self.c_vars:
            self.groups['c'] = G_['c']
        # step sizes live on the device so that captured graphs can be replayed with a new step count
        self.lr_dev = {k: torch.zeros(1, dtype=torch.float32, device=dev) for k in self.groups}
        self._lr_ring = torch.zeros(4096, dtype=torch.float32).pin_memory()
        self._lr_pos = 0
        self._graphs = {}
        self.reducers = {}
        if self.world_size > 1 and not SPLIT_GRAPH:
            from .parallel import GradReducer
            self.reducers = {'d_step': GradReducer(self.d_prog, self.store, ('d',), self.world_size),
                             'g_step': GradReducer(self.g_prog, self.store, ('g', 'c'), self.world_size)}
        self._host_losses = {p.name: torch.zeros(max(len(p.loss_names), 1), dtype=torch.float32).pin_memory()
                             for p in (self.d_prog, self.g_prog)}
        self.counter = 0

    # ------------------------------------------------------------------ train step
    def _push_lr(self, key, value):
        i = self._lr_pos
        self._lr_pos = (i + 1) % self._lr_ring.numel()
        self._lr_ring[i] = value
        self.lr_dev[key].copy_(self._lr_ring[i:i + 1], non_blocking=True)

    def _allreduce(self, group):
        from .parallel import allreduce_sum_
        allreduce_sum_(group.grads, self.world_size)

    def _d_body_a(self):
        g = self.groups['d']
        _C.call('rcgan_zero', g.grads.data_ptr(), g.numel * 4, _C.stream_ptr())
        self.d_prog.run_forward()
        self.d_prog.run_backward()

    def _d_body_b(self):
        cfg = self.config
        adam_step(self.groups['d'], self.lr_dev['d'], cfg.beta1, 0.999, 1e-8, 1.0 / self.world_size)
        self.d_prog.run_updates()

    def _g_body_a(self):
        for k in ('g', 'c'):
            if k in self.groups:
                g = self.groups[k]
                _C.call('rcgan_zero', g.grads.data_ptr(), g.numel * 4, _C.stream_ptr())
        self.g_prog.run_forward()
        self.g_prog.run_backward()

    def _g_body_b(self):
        cfg = self.config
        adam_step(self.groups['g'], self.lr_dev['g'], cfg.beta1, 0.999, 1e-8, 1.0 / self.world_size)
        if 'c' in self.groups:
            adam_step(self.groups['c'], self.lr_dev['c'], cfg.beta1, 0.999, 1e-8, 1.0 / self.world_size)
        self.g_prog.run_updates()

    def _run(self, name, fn):
        """Replay (capturing on first use) the CUDA graph of one launch sequence."""
        if not self.use_cuda_graph:
            fn()
            return
        g = self._graphs.get(name)
        if g is None:
            # warm-up outside capture on a side stream (first-use lazy initialisation), then capture.
            # State mutated by the warm-up run (moving stats, u, Adam moments) is restored afterwards.
            snap = self._snapshot()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                fn()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            self._restore(snap)
            g = torch.cuda.CUDAGraph()
            # (thread-local capture mode: NCCL's watchdog thread keeps polling its events while collectives are captured)
            with torch.cuda.graph(g, capture_error_mode='thread_local' if self.world_size > 1 else 'global'):
                fn()
            self._restore(snap)
            self._graphs[name] = g
        g.replay()

    def _all_vars(self):
        v = dict(self.store.vars)
        if getattr(self, 'r_store', None) is not None:
            v.update(self.r_store.vars)
        return v

    def _snapshot(self):
        snap = {n: v.data.clone() for n, v in self._all_vars().items()}
        for k, g in self.groups.items():
            snap['__m_' + k], snap['__v_' + k] = g.m.clone(), g.v.clone()
        return snap

    def _restore(self, snap):
        for n, v in self._all_vars().items():
            v.data.copy_(snap[n])
        for k, g in self.groups.items():
            g.m.copy_(snap['__m_' + k]); g.v.copy_(snap['__v_' + k])

    def d_step(self):
        cfg = self.config
        g = self.groups['d']
        g.t += 1
        self._push_lr('d', tf_adam_lr(cfg.learning_rate, cfg.beta1, 0.999, g.t))
        if self.world_size > 1 and SPLIT_GRAPH:
            self._run('d_a', self._d_body_a); self._allreduce(g); self._run('d_b', self._d_body_b)
        else:
            # one graph per step; the gradient buckets are all-reduced from inside the backward sweep (parallel.GradReducer)
            red = self.reducers.get('d_step')
            self._run('d', lambda: (self._d_body_a(), red.wait() if red is not None else None, self._d_body_b()))

    def g_step(self):
        cfg = self.config
        g = self.groups['g']
        g.t += 1
        self._push_lr('g', tf_adam_lr(cfg.learning_rate, cfg.beta1, 0.999, g.t))
        if 'c' in self.groups:
            c = self.groups['c']
            c.t += 1
            self._push_lr('c', tf_adam_lr(cfg.learning_rate * cfg.confuse_multiplier, cfg.beta1, 0.999, c.t))
        if self.world_size > 1 and SPLIT_GRAPH:
            self._run('g_a', self._g_body_a)
            self._allreduce(g)
            if 'c' in self.groups:
                self._allreduce(self.groups['c'])
            self._run('g_b', self._g_body_b)
        else:
            red = self.reducers.get('g_step')
            self._run('g', lambda: (self._g_body_a(), red.wait() if red is not None else None, self._g_body_b()))

    def feed(self, batch_images=None, batch_z=None, batch_labels_real=None, batch_labels_gen=None, batch_labels_fake=None,
             batch_labels_real_weights=None):
        """feed_dict equivalent: host (ideally pinned) or device tensors -> the programs' fp32 input buffers."""
        def put(prog, name, src):
            if src is None or name not in prog.inputs:
                return
            t = prog.inputs[name]
            src = torch.as_tensor(src)
            if src.dtype != torch.float32:
                src = src.to(torch.float32)
            t.data.copy_(src.reshape(-1), non_blocking=True)
        for prog in (self.d_prog, self.g_prog):
            put(prog, 'inputs', batch_images); put(prog, 'z', batch_z); put(prog, 'y_real', batch_labels_real)
            put(prog, 'y_gen', batch_labels_gen); put(prog, 'y_fake', batch_labels_fake)
            put(prog, 'y_real_weights', batch_labels_real_weights)

    def train_iteration(self, fetch_losses=True, **feeds):
        """One iteration of the reference hot loop (mnist/model.py:337-372): D step, then 2 x (G step + C step)
        with the same z and labels.  Returns the loss dict (device->host read) when fetch_losses."""
        if feeds:
            self.feed(**feeds)
        self.d_step()
        if fetch_losses:
            self._host_losses['d_step'].copy_(self.d_prog.losses, non_blocking=True)
        self.g_step()
        self.g_step()
        self.counter += 1
        if not fetch_losses:
            return None
        self._host_losses['g_step'].copy_(self.g_prog.losses, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        out = self.d_prog.loss_dict(self._host_losses['d_step'])
        out.update(self.g_prog.loss_dict(self._host_losses['g_step']))
        out['d_loss'] = out.get('d_loss_real', 0.0) + out.get('d_loss_fake', 0.0)
        if 'class_loss_fake' in out:
            pass
        return out

    def sample(self, z, y_gen):
        """self.sampler evaluation (gen_sampler)."""
        self.s_prog.inputs['z'].data.copy_(torch.as_tensor(z, dtype=torch.float32).reshape(-1), non_blocking=True)
        self.s_prog.inputs['y_gen'].data.copy_(torch.as_tensor(y_gen, dtype=torch.float32).reshape(-1), non_blocking=True)
        self.s_prog.run_forward()
        return self.sampler.torch().clone()

    # ------------------------------------------------------------------ label recovery (mnist/model.py:494-640)
    def build_recover(self, recover_batch_size, seed=0):
        """Graph of DCGAN.recover_labels for R = recover_batch_size real images: variables z_recover [R*y_dim, z_dim] and
        y_logit_recover [R, y_dim] (tf.get_variable default = Glorot uniform), gen_sampler (batch norm in inference mode) on
        the R*y_dim (z, one-hot) pairs, mse_loss (:538-541), plain gradient descent on the two variables (:612-617)."""
        R, k, dev = int(recover_batch_size), self.y_dim, self.device
        self.recover_R = R
        self.r_store = rs = VariableStore(dev, lambda name: 'r')
        gen = torch.Generator().manual_seed(seed)

        def glorot(shape):
            lim = math.sqrt(6.0 / (shape[0] + shape[1]))
            return (torch.rand(tuple(shape), generator=gen, dtype=torch.float32) * 2 - 1) * lim
        self.y_logit_recover = rs.get('y_logit_recover', (R, k), glorot)
        self.z_recover = rs.get('z_recover', (R * k, self.z_dim), glorot)
        rs.finalize()
        self.r_prog = rp = Program('recover', dev, self.act_dtype)
        saved_B = self.batch_size
        self.batch_size = R * k                       # the reference re-purposes self.batch_size the same way (:526)
        try:
            with rp:
                actual = rp.input('sample_actual', [R, self.output_height * self.output_width * self.c_dim])
                hard_y = rp.input('hard_y_recover', [R * k, k])
                hard_y.data.copy_(torch.eye(k, device=dev).repeat(R, 1).reshape(-1))
                self.y_recover = SoftmaxRowsOp(self.y_logit_recover).y
                img = self.gen_sampler(CastOp(self.z_recover, self.act_dtype).y, hard_y)
                img32 = CastOp(img, _C.F32).y
                self.sample_recover_each_y = img32.view([R * k, self.output_height * self.output_width * self.c_dim])
                self.recover_loss_op = RecoverMSEOp(self.sample_recover_each_y, actual, self.y_recover, 'mse_loss')
        finally:
            self.batch_size = saved_B
        rp.finalize([self.z_recover, self.y_logit_recover])
        self._r_host = torch.zeros(1, dtype=torch.float32).pin_memory()
        # graphs captured over a previous recover program replay ITS buffers: drop them (a second build_recover /
        # recover_labels call starts from freshly initialised z_recover / y_logit_recover, like the reference's re-initialisation)
        for key in [k_ for k_ in self._graphs if k_.startswith('recover_')]:
            del self._graphs[key]
        return rp

    def _r_body(self, lr):
        g = self.r_store.groups['r']
        _C.call('rcgan_zero', g.grads.data_ptr(), g.numel * 4, _C.stream_ptr())
        self.r_prog.run_forward()
        self.r_prog.run_backward()
        _C.call('rcgan_sgd', g.params.data_ptr(), g.grads.data_ptr(), g.numel, float(lr), 1.0, _C.stream_ptr())

    def recover_step(self, sample_actual=None, learning_rate=500.0, fetch=True):
        """One `sess.run(recover_optim)` (:626-629): returns mse_loss evaluated BEFORE the update, like the reference's fetch."""
        if sample_actual is not None:
            src = torch.as_tensor(sample_actual, dtype=torch.float32)
            self.r_prog.inputs['sample_actual'].data.copy_(src.reshape(-1), non_blocking=True)
        self._run('recover_%g' % learning_rate, lambda: self._r_body(learning_rate))
        if not fetch:
            return None
        self._r_host.copy_(self.r_prog.losses, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(self._r_host[0])

    def recover_labels(self, sample_actual, y_actual=None, recover_epoch=1000, learning_rate=500.0, log_every=100):
        """The reference's loop (:622-640) on a given batch of real images; returns (y_recover [R,k], mse_loss,
        zero_one_loss = tf.losses.cosine_distance(y_actual, one_hot(argmax y_recover)) = 1 - accuracy).
        Called as the reference does, `recover_labels(config)` (mnist/main.py:142, model.py:494-640), it restores the latest
        checkpoint (:497-502), draws recover_batch_size random rows of self.data_X from the numpy stream (:613-617) and runs
        recover_epoch steps at recover_learning_rate."""
        if hasattr(sample_actual, 'recover_batch_size'):
            config = sample_actual
            could_load, _ = self.load(self.checkpoint_dir)
            if not could_load:
                raise Exception(" [!] Load failed...")
            idx = np.random.randint(len(self.data_X), size=[config.recover_batch_size])
            y_rec, mse, zo = self.recover_labels(self.data_X[idx], self.data_y_actual[idx], config.recover_epoch,
                                                 config.recover_learning_rate, log_every)
            print("Recover: mse_loss: %.5g, zeroone_loss: %.5g" % (mse, zo))
            return y_rec, mse, zo
        if not hasattr(self, 'r_prog') or self.recover_R != len(sample_actual):
            self.build_recover(len(sample_actual))
        mse = None
        for epoch in range(recover_epoch):
            mse = self.recover_step(sample_actual if epoch == 0 else None, learning_rate,
                                    fetch=(epoch + 1) % log_every == 0 or epoch == recover_epoch - 1)
        self.r_prog.run_forward()
        y_rec = self.y_recover.torch().clone().cpu()
        zero_one = None
        if y_actual is not None:
            ya = torch.as_tensor(y_actual, dtype=torch.float32)
            onehot = torch.eye(self.y_dim)[y_rec.argmax(-1)]
            zero_one = float((1.0 - (ya * onehot).sum(-1)).mean())
        return y_rec, mse, zero_one

    # ------------------------------------------------------------------ checkpoints (mnist/model.py:836-867)
    @property
    def model_dir(self):
        return "{}_{}_{}_{}".format(self.dataset_name, self.batch_size, self.output_height, self.output_width)

    def save(self, checkpoint_dir, step):
        """:842-851: every global variable under its TF name + the Adam slots, `DCGAN.model-<step>` (checkpoint.py)"""
        from . import checkpoint
        return checkpoint.save(self, os.path.join(checkpoint_dir, self.model_dir), "DCGAN.model", step)

    def load(self, checkpoint_dir):
        """:853-867 -> (could_load, counter)"""
        from . import checkpoint
        print(" [*] Reading checkpoints...")
        if checkpoint_dir is None:
            return False, 0
        path = checkpoint.latest_checkpoint(os.path.join(checkpoint_dir, self.model_dir))
        if path and os.path.exists(path + '.npz'):
            checkpoint.restore(self, path)      # (captured graphs read parameters / Adam state from fixed addresses: no re-capture)
            print(" [*] Success to read {}".format(os.path.basename(path)))
            return True, checkpoint.step_of(path)
        print(" [*] Failed to find a checkpoint")
        return False, 0

    def noise_schedule(self, epoch):
        """mnist/model.py:293-321 (--add_noise, run_rcgany.sh): the one-coin confusion matrix the labels are re-noised with at
        the start of `epoch`.  Returns (noise_alpha, noise_C)."""
        cfg, k = self.config, self.y_dim
        alpha_start = (self.noise_alpha - (1. - self.alpha) / (k - 1)) / (self.alpha - (1. - self.alpha) / (k - 1))
        alpha_start = min(1.0, alpha_start)
        if self.noise_alpha > 0.9:
            raise ValueError('same rate activated, but effective noise alpha {} > 0.9!'.format(self.noise_alpha))
        if alpha_start == 1.:
            end_epoch = cfg.noise_start
        else:
            end_epoch = cfg.noise_start + ((cfg.noise_end - cfg.noise_start) / (0.9 - self.noise_alpha) * (self.alpha - self.noise_alpha))
            end_epoch = min(cfg.noise_end, end_epoch)
        if epoch < cfg.noise_start:
            noise_alpha = alpha_start
        elif epoch < end_epoch:
            noise_alpha = alpha_start + (1. - alpha_start) * (epoch - cfg.noise_start) / (end_epoch - cfg.noise_start)
        else:
            noise_alpha = 1.0
        noise_alpha = min(1.0, noise_alpha)
        return noise_alpha, one_coin_confusion(noise_alpha, k)

    def renoise_labels(self, epoch):
        """mnist/model.py:323-333: data_y_real / data_y_fake are re-drawn through noise_C, sample by sample (real then fake), from
        the same numpy stream -- on the device, bit-exact (rcgan_sample_renoise_mnist).  Cumulative, as in the reference: the
        labels of the previous epoch are the input."""
        _, noise_C = self.noise_schedule(epoch)
        real, fake = self.sampler_state.renoise_mnist(self.data_y_real.argmax(-1), self.data_y_fake.argmax(-1), noise_C)
        eye = np.eye(self.y_dim)
        self.data_y_real, self.data_y_fake = eye[real], eye[fake]

    def train(self, config=None, max_iters=None, log_every=100):
        """mnist/model.py:249-491 hot loop over self.data_*: the sample_z draw (:274), per epoch the --add_noise re-noising of
        the real / fake labels (:293-333), per batch the z draw (:342) and 1 D + 2 G(+C) steps (:344-372) -- every random number
        from the one numpy-legacy stream seeded in load_mnist, so batch_z and the labels follow the reference's stream position.
        (Logging evals, checkpoints every 500 steps (`save`), sample grids and the frozen-classifier metric are outside the hot
        path; see checkpoint.py / metrics.py.)"""
        config = config or self.config
        n = len(self.data_X)
        batch_idxs = min(n, config.train_size) // config.batch_size
        B = config.batch_size
        start_time = time.time()
        pin = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).pin_memory()
        X, yg, yw = pin(self.data_X), pin(self.data_y_gen), pin(self.data_y_real_weights)
        yr, yf = pin(self.data_y_real), pin(self.data_y_fake)
        self.sample_z = self.sampler_state.uniform(-1, 1, self.sample_num * self.z_dim).reshape(self.sample_num, self.z_dim)
        it = 0
        counter = 1
        could_load, checkpoint_counter = self.load(self.checkpoint_dir)          # :284-290
        if could_load:
            counter = checkpoint_counter
            print(" [*] Load SUCCESS")
        else:
            print(" [!] Load failed...")
        for epoch in range(config.epoch):
            if self.add_noise:
                self.renoise_labels(epoch)
                yr, yf = pin(self.data_y_real), pin(self.data_y_fake)
            for idx in range(0, int(batch_idxs)):
                sl = slice(idx * B, (idx + 1) * B)
                batch_z = self.sampler_state.uniform(-1, 1, B * self.z_dim).reshape(B, self.z_dim)
                losses = self.train_iteration(fetch_losses=(it % log_every == 0), batch_images=X[sl], batch_z=batch_z,
                                              batch_labels_real=yr[sl], batch_labels_gen=yg[sl], batch_labels_fake=yf[sl],
                                              batch_labels_real_weights=yw[sl])
                if losses is not None:
                    print("Epoch: [%2d] [%4d/%4d] time: %4.4f, d_loss: %.8f, g_loss: %.8f" % (
                        epoch, idx, batch_idxs, time.time() - start_time, losses['d_loss'], losses['g_loss']))
                it += 1
                counter += 1
                if self.checkpoint_dir is not None and np.mod(counter, 500) == 2:     # :489-490
                    self.save(self.checkpoint_dir, counter)
                if max_iters is not None and it >= max_iters:
                    if self.checkpoint_dir is not None:
                        self.save(self.checkpoint_dir, counter)
                    return
