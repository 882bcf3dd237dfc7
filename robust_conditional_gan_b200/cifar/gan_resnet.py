"""Drop-in for the hot path of cifar10/gan_resnet.py: the SN-ResNet generator / projection discriminator
(:199-421), the permutation classifier (:458-483), the RCGAN loss graph (:498-786) and one training iteration
(:919-947), recorded as static programs of librcgan_b200.so launches.

Towers (`DEVICES`, gan_resnet.py:183-192) are ranks here: each process holds one tower's shard (n = BATCH_SIZE / T),
computes its own conditional-BN statistics (the reference does per tower) and the gradient is the NCCL mean over ranks.
"""
import math
from types import SimpleNamespace

import numpy as np
import torch

from .. import _C, scope as S
from ..graph import I32, U8, Program, VariableStore, tf_adam_lr
from ..nnops import (ActOp, AddOp, CastOp, ChannelLossOp, ConcatRowsOp, ConvOp, GatherRowsOp, MeanHWOp, Pool2Op, PreprocessCifarOp, RandomFillOp,
                     SigmoidCEOp,
                     SoftmaxRowsOp, Upsample2Op, adam_step)
from . import ops as lib_ops

NO_OPS = 'NO_OPS'
FUSE_RESIDUAL = __import__('os').environ.get('RCGAN_FUSE_RESIDUAL', '1') == '1'
# ... in the discriminator too (its 128-channel convs are bound by the operand stream and their epilogue is NOT hidden: measured,
# tools/epi_bench.py, the fused residual read costs +56 us on the 16x16 layer against ~22 us for the separate add kernel)
FUSE_RESIDUAL_D = __import__('os').environ.get('RCGAN_FUSE_RESIDUAL_D', '0') == '1'
FUSE_RELU_OUT_D = __import__('os').environ.get('RCGAN_FUSE_RELU_OUT_D', '1') == '1'
# 3x3 ConvMeanPool / UpsampleConv as ONE 4x4 stride-2 conv / conv2d_transpose with the folded filter; 0 = A/B switch
FOLD_RESAMPLE = __import__('os').environ.get('RCGAN_FOLD', '1') == '1'
SPLIT_GRAPH = __import__('os').environ.get('RCGAN_DP_SPLIT_GRAPH', '0') == '1'
Z_DIM, VOCAB_SIZE, EMBEDDING_DIM, IMG_SIZE, IMG_DIM, OUTPUT_DIM = 128, 10, 300, 32, 3, 3072
N_CRITIC, GEN_BS_MULTIPLE = 5, 2


def default_flags(**kw):
    """cifar10/gan_resnet.py:40-76 (hot-path subset)."""
    f = SimpleNamespace(dataset='cifar', algorithm='rcgan', alpha=0.8, batch_size=64, niters=50000, lr=2e-4, ngpus=2,
                        multi_gpu_multi_batch=True, confuse_init=False, confuse_init_diag=0.2, confuse_multiplier=1.0,
                        confuse_lr_decay=False, perm_classifier=False, perm_multiplier=1.0, perm_type='linear')
    for k, v in kw.items():
        setattr(f, k, v)
    return f


class Net:
    """The module-level constants + network functions of gan_resnet.py, bound to one configuration."""

    def __init__(self, algorithm, dim_g=128, dim_d=128, towers=1, perm_type='linear'):
        self.ALGORITHM, self.DIM_G, self.DIM_D = algorithm, dim_g, dim_d
        # towers > 1: the reference's in-process towers (gan_resnet.py:183-192, 529-546) -- the batch of this process is `towers`
        # equal sample ranges; only the generator's conditional batch norms see the difference (statistics per tower)
        self.towers, self.perm_type = towers, perm_type

    # gan_resnet.py:199-205
    @staticmethod
    def nonlinearity(x, activation_fn='relu', leakiness=0.2):
        return ActOp(x, 'relu' if activation_fn == 'relu' else 'lrelu', leakiness).y

    def Normalize(self, name, inputs, labels=None, fuse_act=None):
        """:207-228 with the shipped constants (CONDITIONAL, NORMALIZATION_G, not NORMALIZATION_D, not ACGAN):
        conditional batch norm in G, identity in D."""
        with S.variable_scope(name):
            if 'G.' in name and labels is not None:
                return lib_ops.cond_batchnorm(name, [0, 1, 2], inputs, labels=labels, n_labels=10, fuse_act=fuse_act,
                                              groups=self.towers)
        if fuse_act == 'relu':
            # D: the bare nonlinearity(inputs) (:318).  When `inputs` is the output of a conv with a fused-epilogue path (the previous
            # block's Conv2 + shortcut), that conv's epilogue writes relu(inputs) as a second output: no separate pass
            prod = getattr(inputs.base, 'producer', None)
            if FUSE_RELU_OUT_D and isinstance(prod, ConvOp) and prod.y.base is inputs.base and prod.can_emit_relu():
                return prod.emit_relu()
        return ActOp(inputs, fuse_act).y if fuse_act else inputs

    @staticmethod
    def ConvMeanPool(inputs, output_dim, filter_size=3, stride=1, name=None, spectral_normed=False, update_collection=None,
                     inputs_norm=False, he_init=True, biases=True, residual=None):
        """:231-241 (residual: the block's shortcut, added by the folded conv's epilogue)"""
        if FOLD_RESAMPLE and filter_size == 3 and inputs.shape[1] % 2 == 0 and inputs.shape[2] % 2 == 0:
            return lib_ops.Conv2D(inputs, inputs.shape[-1], output_dim, filter_size, stride, name, spectral_normed=spectral_normed,
                                  update_collection=update_collection, he_init=he_init, biases=biases, fold='pool', residual=residual)
        assert residual is None
        out = lib_ops.Conv2D(inputs, inputs.shape[-1], output_dim, filter_size, stride, name, spectral_normed=spectral_normed,
                             update_collection=update_collection, he_init=he_init, biases=biases)
        return Pool2Op(out).y

    @staticmethod
    def MeanPoolConv(inputs, output_dim, filter_size=3, stride=1, name=None, spectral_normed=False, update_collection=None,
                     inputs_norm=False, he_init=True, biases=True):
        """:244-256"""
        out = Pool2Op(inputs).y
        return lib_ops.Conv2D(out, out.shape[-1], output_dim, filter_size, stride, name, spectral_normed=spectral_normed,
                              update_collection=update_collection, he_init=he_init, biases=biases)

    @staticmethod
    def UpsampleConv(inputs, output_dim, filter_size=3, stride=1, name=None, spectral_normed=False, update_collection=None,
                     inputs_norm=False, he_init=True, biases=True, pre_norm=False):
        """:259-272"""
        if FOLD_RESAMPLE and filter_size == 3:
            return lib_ops.Conv2D(inputs, inputs.shape[-1], output_dim, filter_size, stride, name, spectral_normed=spectral_normed,
                                  update_collection=update_collection, he_init=he_init, biases=biases, pre_norm=pre_norm, fold='up')
        up = Upsample2Op(inputs)
        out = up.y
        return lib_ops.Conv2D(out, out.shape[-1], output_dim, filter_size, stride, name, spectral_normed=spectral_normed,
                              update_collection=update_collection, he_init=he_init, biases=biases, pre_norm=pre_norm, up_op=up)

    def ResidualBlock(self, inputs, input_dim, output_dim, filter_size, name, spectral_normed=False, update_collection=None,
                      inputs_norm=False, resample=None, labels=None, biases=True):
        """:275-328 pre-activation block."""
        kw = dict(spectral_normed=spectral_normed, update_collection=update_collection, biases=biases)
        if resample == 'down':
            conv_1 = lambda x, **k: lib_ops.Conv2D(x, input_dim, input_dim, **k)
            conv_2 = lambda x, **k: self.ConvMeanPool(x, output_dim, **k)
            conv_shortcut = self.ConvMeanPool
        elif resample == 'up':
            conv_1 = lambda x, **k: self.UpsampleConv(x, output_dim, **k)
            conv_shortcut = self.UpsampleConv
            conv_2 = lambda x, **k: lib_ops.Conv2D(x, output_dim, output_dim, **k)
        elif resample is None:
            conv_shortcut = lambda x, output_dim, **k: lib_ops.Conv2D(x, input_dim, output_dim, **k)
            conv_1 = lambda x, **k: lib_ops.Conv2D(x, input_dim, output_dim, **k)
            conv_2 = lambda x, **k: lib_ops.Conv2D(x, output_dim, output_dim, **k)
        else:
            raise Exception('invalid resample value')
        if output_dim == input_dim and resample is None:
            shortcut = inputs
        else:
            # 1x1 shortcuts commute with the resampling (a 1x1 conv touches one pixel; nearest-neighbour upsampling copies
            # pixels, mean-pooling averages them and the conv is linear): UpsampleConv_1x1(x) == Upsample(Conv_1x1(x)) bit for
            # bit, ConvMeanPool_1x1(x) == Conv_1x1(MeanPool(x)) up to fp32 summation order.  Running the conv on the SMALL
            # grid is a 4x flop / traffic cut (legal for parity; rooflines still use the reference flops).
            if resample == 'up':
                shortcut = lib_ops.Conv2D(inputs, input_dim, output_dim, 1, 1, name + '.Shortcut', he_init=False, **kw)
                if not (FUSE_RESIDUAL and ('G.' in name or FUSE_RESIDUAL_D)):
                    shortcut = Upsample2Op(shortcut).y
            elif resample == 'down':
                shortcut = lib_ops.Conv2D(Pool2Op(inputs).y, input_dim, output_dim, 1, 1, name + '.Shortcut', he_init=False, **kw)
            else:
                shortcut = conv_shortcut(inputs, output_dim=output_dim, filter_size=1, name=name + '.Shortcut', he_init=False, **kw)
        output = self.Normalize(name + '.N1', inputs, labels=labels, fuse_act='relu')
        if 'G.' in name and labels is not None:
            output = conv_1(output, filter_size=filter_size, name=name + '.Conv1', he_init=True, **kw)
            output = self.Normalize(name + '.N2', output, labels=labels, fuse_act='relu')
        else:
            # D: Normalize is the identity (NORMALIZATION_D = False), so N2 + nonlinearity is a bare relu on Conv1's output:
            # it rides in Conv1's epilogue instead of a separate pass
            output = conv_1(output, filter_size=filter_size, name=name + '.Conv1', he_init=True, fuse_act='relu', **kw)
        fuse = FUSE_RESIDUAL and ('G.' in name or FUSE_RESIDUAL_D)
        if not fuse or (resample == 'down' and not FOLD_RESAMPLE):
            output = conv_2(output, filter_size=filter_size, name=name + '.Conv2', he_init=True, **kw)
            return AddOp(shortcut, output).y
        # `shortcut + output` (:328) in Conv2's epilogue (rcgan_conv2d_fprop_ex): bit-identical to the separate add; an 'up' block's
        # shortcut stays at half resolution and is upsampled by the epilogue's index map
        if resample == 'up':
            return conv_2(output, filter_size=filter_size, name=name + '.Conv2', he_init=True, residual=shortcut, residual_up=True, **kw)
        return conv_2(output, filter_size=filter_size, name=name + '.Conv2', he_init=True, residual=shortcut, **kw)

    def OptimizedResBlockDisc1(self, inputs, spectral_normed=False, update_collection=None, inputs_norm=False, biases=True):
        """:331-353"""
        kw = dict(spectral_normed=spectral_normed, update_collection=update_collection, biases=biases)
        shortcut = self.MeanPoolConv(inputs, output_dim=self.DIM_D, filter_size=1, name='D.Block.1.Shortcut', he_init=False, **kw)
        output = lib_ops.Conv2D(inputs, IMG_DIM, self.DIM_D, 3, 1, 'D.Block.1.Conv1', he_init=True, fuse_act='relu', **kw)
        if FUSE_RESIDUAL and FUSE_RESIDUAL_D and FOLD_RESAMPLE:
            return self.ConvMeanPool(output, self.DIM_D, filter_size=3, name='D.Block.1.Conv2', he_init=True, residual=shortcut, **kw)
        output = self.ConvMeanPool(output, self.DIM_D, filter_size=3, name='D.Block.1.Conv2', he_init=True, **kw)
        return AddOp(shortcut, output).y

    def Generator(self, n_samples_, labels, noise=None, reuse=False):
        """:356-371.  noise: graph tensor [n,128] (the in-graph tf.random_normal is replaced by a fed tensor).
        Returns the NHWC image tensor [n,32,32,3] (the reference flattens it to [n,3072])."""
        D = self.DIM_G
        with S.variable_scope("Generator"):
            output = lib_ops.Linear(noise, 128, 4 * 4 * D * 8, 'G.Input')
            output = output.view([n_samples_, 4, 4, D * 8])
            output = self.ResidualBlock(output, D * 8, D * 2, 3, 'G.Block.1', resample='up', labels=labels, biases=True)
            output = self.ResidualBlock(output, D * 2, D * 2, 3, 'G.Block.2', resample='up', labels=labels, biases=True)
            output = self.ResidualBlock(output, D * 2, D * 2, 3, 'G.Block.3', resample='up', labels=labels, biases=True)
            output = self.Normalize('G.OutputNorm', output, labels, fuse_act='relu')
            return lib_ops.Conv2D(output, D * 2, IMG_DIM, 3, 1, 'G.Output', he_init=False, fuse_act='tanh')

    def Discriminator(self, inputs, labels, update_collection=None, reuse=False):
        """:374-412.  Returns (output [N, DIM_D], output_wgan [N,1])."""
        D = self.DIM_D
        with S.variable_scope("Discriminator"):
            output = self.OptimizedResBlockDisc1(inputs, spectral_normed=True, update_collection=update_collection, biases=True)
            output = self.ResidualBlock(output, D, D, 3, 'D.Block.2', spectral_normed=True, update_collection=update_collection,
                                        resample='down', labels=None, biases=True)
            for i in (3, 4, 5, 6):
                output = self.ResidualBlock(output, D, D, 3, 'D.Block.%d' % i, spectral_normed=True,
                                            update_collection=update_collection, resample=None, labels=None, biases=True)
            output = MeanHWOp(output, relu=True).y                 # nonlinearity + reduce_mean(axis=[1,2])
            out32 = CastOp(output, _C.F32).y
            output_wgan = lib_ops.Linear(out32, D, 1, 'D.Output', spectral_normed=True, update_collection=update_collection)
            return out32, output_wgan

    def Discriminator_projection(self, labels, update_collection=None, reuse=False):
        """:414-421.  labels None -> the embeddings of all VOCAB_SIZE classes [10, DIM_D]."""
        with S.variable_scope("Discriminator"):
            embedding_y = lib_ops.embed_y(labels, VOCAB_SIZE, EMBEDDING_DIM)
            return lib_ops.Linear(embedding_y, EMBEDDING_DIM, self.DIM_D, 'D.Embedding_y', spectral_normed=True,
                                  update_collection=update_collection, biases=True)

    def perm_classifier(self, x, reuse=False):
        """:456-483; x fp32 [n,32,32,3].  'linear': one SN-Linear 3072 -> 10; '2layer': SN-Linear 3072 -> 128 -> 10 (no
        nonlinearity in between, as in the reference)."""
        with S.variable_scope("Discriminator"):
            flat = x.view([x.shape[0], OUTPUT_DIM])
            if self.perm_type == 'linear':
                return lib_ops.Linear(flat, OUTPUT_DIM, VOCAB_SIZE, 'D.d_perm_classifier_h1', spectral_normed=True, biases=True)
            if self.perm_type == '2layer':
                hidden = lib_ops.Linear(flat, OUTPUT_DIM, 128, 'D.d_perm_classifier_h1', spectral_normed=True, biases=True)
                return lib_ops.Linear(hidden, 128, VOCAB_SIZE, 'D.d_perm_classifier_h2', spectral_normed=True, biases=True)
            raise ValueError('Unknown perm_type {}'.format(self.perm_type))


def _group_of(name):
    """gan_resnet.py:788-800: optimizer ownership by substring."""
    if name == 'confusion_logits':
        return 'c'
    if 'Discriminator' in name:
        return 'd'
    if 'Generator' in name:
        return 'g'
    return None


def lr_decay(it):
    """:700-703"""
    return max(0., 1. - it / 100000.) if it < 50000 else 0.5


class RCGANCifar(object):
    """The graph of gan_resnet.main() (:498-817) for ONE tower and its training iteration (:919-947)."""

    def __init__(self, flags=None, tower_batch=32, device='cuda', precision='bf16', seed=0, world_size=1, rank=0, dim=128,
                 use_cuda_graph=True, rng='fed', towers=1):
        """rng: 'fed' -- the generator noise and the dequantisation noise are program inputs (parity tests, the benchmark's
        host-fed leg); 'device' -- they are drawn in-graph like the reference's tf.random_normal / tf.random_uniform
        (gan_resnet.py:363-364, 550) by a Philox kernel keyed by (seed, rank, step): nothing but pixels and labels crosses PCIe."""
        self.FLAGS = flags if flags is not None else default_flags()
        # tower_batch = BATCH_SIZE / len(DEVICES); towers = towers of THIS process (1: towers are ranks; 2 reproduces the
        # reference's single-GPU graph, DEVICES = [gpu0, gpu0]: conditional-BN statistics over each half of the batch)
        self.towers = int(towers)
        self.n = tower_batch * self.towers
        self.device = torch.device(device)
        self.act_dtype = {'bf16': _C.BF16, 'fp32': _C.F32}[precision]
        self.world_size, self.rank, self.seed, self.dim = world_size, rank, seed, dim
        self.use_cuda_graph = use_cuda_graph
        self.rng = rng
        self.rng_step = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._step_ring = torch.zeros(1024, dtype=torch.int64).pin_memory()
        self._step_count = 0
        if self.FLAGS.algorithm not in ('rcgan', 'rcgan-u', 'biased', 'unbiased'):
            raise ValueError('unknown algorithm ' + str(self.FLAGS.algorithm))
        if getattr(self.FLAGS, 'perm_type', 'linear') not in ('linear', '2layer'):
            raise ValueError('Unknown perm_type {}'.format(self.FLAGS.perm_type))
        if not _C.load().rcgan_device_ok():
            raise _C.RcganError('rcgan_b200 needs an sm_100 device: ' + _C.last_error())
        a = self.FLAGS.alpha
        self.C_ALPHA = ((1 - a) / 9.0) * np.ones((10, 10)) + (a - (1 - a) / 9.0) * np.eye(10)   # :106
        self.build()

    # ------------------------------------------------------------------ graph
    def _confusion(self, prog):
        """:498-524"""
        F = self.FLAGS
        if F.algorithm == 'rcgan-u':
            def init(shape):
                if F.confuse_init:
                    aa = 7.0 if F.confuse_init_diag > 0.99 else np.log(VOCAB_SIZE * F.confuse_init_diag / (1. - F.confuse_init_diag))
                    aa = min(7.0, aa)
                    ci = (0 - aa / VOCAB_SIZE) * np.ones([VOCAB_SIZE, VOCAB_SIZE], dtype=np.float32)
                    np.fill_diagonal(ci, aa - aa / VOCAB_SIZE)
                    return torch.as_tensor(ci)
                lim = math.sqrt(6.0 / (2 * VOCAB_SIZE))
                return (torch.rand(tuple(shape), generator=S.init_generator()) * 2 - 1) * lim
            return SoftmaxRowsOp(S.get_variable('confusion_logits', [VOCAB_SIZE, VOCAB_SIZE], init)).y
        C = prog.new((VOCAB_SIZE, VOCAB_SIZE), _C.F32)
        C.data.copy_(torch.as_tensor(self.C_ALPHA, dtype=torch.float32).reshape(-1))
        return C

    def _eye(self, prog):
        if not hasattr(prog, '_eye'):
            prog._eye = prog.new((VOCAB_SIZE, VOCAB_SIZE), _C.F32)
            prog._eye.data.copy_(torch.eye(VOCAB_SIZE, device=self.device).reshape(-1))
        return prog._eye

    def build(self):
        F, n, dev = self.FLAGS, self.n, self.device
        alg = F.algorithm
        net = self.net = Net(alg, self.dim, self.dim, self.towers, getattr(F, 'perm_type', 'linear'))
        self.store = VariableStore(dev, _group_of)
        S.set_store(self.store, self.seed)
        onehot = lambda prog, lab: GatherRowsOp(self._eye(prog), lab).out
        # ---------------- D step (:526-697)
        self.d_prog = dp_ = Program('d_step', dev, self.act_dtype)
        with dp_:
            raw = dp_.input('all_real_data_int', [n, OUTPUT_DIM], U8)           # pixels travel as bytes (the reference: int32)
            labels = dp_.input('all_real_labels', [n, 1], I32)
            labels_random = dp_.input('all_random_labels', [n, 1], I32)
            labels_biased = dp_.input('all_labels_biased', [n, 1], I32)
            inv_w = dp_.input('all_labels_inv_weights', [n, VOCAB_SIZE])
            rseed = (self.seed * 1000003 + self.rank) & 0xffffffffffff
            if self.rng == 'device':
                noise = RandomFillOp([n, Z_DIM], True, 0.0, 1.0, rseed, self.rng_step).y
                dq = RandomFillOp([n, OUTPUT_DIM], False, 0.0, 1.0 / 128, rseed, self.rng_step).y
            else:
                noise = dp_.input('noise', [n, Z_DIM])
                dq = dp_.input('dequant_noise', [n, OUTPUT_DIM])       # tf.random_uniform(0, 1/128), fed (zeros for parity)
            real32 = PreprocessCifarOp(raw, dq, _C.F32).y
            real = CastOp(real32, self.act_dtype).y
            fake = net.Generator(n, labels_random, noise=CastOp(noise, self.act_dtype).y)
            # one D pass over [real; fake] as the reference does (:558-583; D has no batch statistics, so the halves are
            # independent): every kernel of the trunk sees 2n samples instead of two launches of n
            self.fake_D = fake
            h_all, psi_all = net.Discriminator(ConcatRowsOp(real, fake).y, None, update_collection=None)
            self.h_all = h_all
            V = net.Discriminator_projection(None, update_collection=None)
            w_real = inv_w if alg == 'unbiased' else onehot(dp_, labels)
            if alg == 'rcgan-u':
                w_fake = GatherRowsOp(self._confusion(dp_), labels_random).out
            elif alg == 'rcgan':
                w_fake = onehot(dp_, labels_biased)
            else:
                w_fake = onehot(dp_, labels_random)
            self.disc_real = ChannelLossOp(h_all, psi_all, V, w_real, _C.HINGE_D_REAL, 'disc_real_l', rows=(0, n)).logits
            self.disc_fake = ChannelLossOp(h_all, psi_all, V, w_fake, _C.HINGE_D_FAKE, 'disc_fake_l', rows=(n, n)).logits
            if F.perm_classifier:
                SigmoidCEOp(net.perm_classifier(real32), onehot(dp_, labels), 'perm_classifier_real_loss', 1.0)
        # ---------------- G step (:715-786)
        self.g_prog = gp_ = Program('g_step', dev, self.act_dtype)
        with gp_:
            m = GEN_BS_MULTIPLE * n
            noise = (RandomFillOp([m, Z_DIM], True, 0.0, 1.0, rseed, self.rng_step).y if self.rng == 'device'
                     else gp_.input('noise', [m, Z_DIM]))
            lab_rand = gp_.input('all_random_labels_G', [m, 1], I32)
            lab_bias = gp_.input('all_labels_biased_G', [m, 1], I32)
            fakeG = net.Generator(m, lab_rand, noise=CastOp(noise, self.act_dtype).y, reuse=True)
            h, psi = net.Discriminator(fakeG, lab_bias, update_collection=NO_OPS, reuse=True)
            V = net.Discriminator_projection(None, update_collection=None, reuse=True)
            if alg == 'rcgan-u':
                w = GatherRowsOp(self._confusion(gp_), lab_rand).out
            elif alg == 'rcgan':
                w = onehot(gp_, lab_bias)
            else:
                w = onehot(gp_, lab_rand)
            ChannelLossOp(h, psi, V, w, _C.HINGE_G, 'gen_wgan')
            if F.perm_classifier:
                SigmoidCEOp(net.perm_classifier(CastOp(fakeG, _C.F32).y, reuse=True), onehot(gp_, lab_rand),
                            'perm_classifier_fake_loss', F.perm_multiplier)
            self.fake_G = fakeG
        self.store.finalize()
        V_ = self.store.vars
        self.disc_params = [v for k, v in V_.items() if v.trainable and 'Discriminator' in k]
        self.gen_params = [v for k, v in V_.items() if v.trainable and 'Generator' in k]
        self.c_params = [V_['confusion_logits']] if 'confusion_logits' in V_ else []
        self.d_prog.finalize(self.disc_params)
        self.g_prog.finalize(self.gen_params + self.c_params)
        self.groups = {k: g for k, g in self.store.groups.items()}
        self.lr_dev = {k: torch.zeros(1, dtype=torch.float32, device=dev) for k in self.groups}
        self._lr_ring = torch.zeros(4096, dtype=torch.float32).pin_memory()
        self._lr_pos = 0
        self._graphs = {}
        self._host_losses = {p.name: torch.zeros(max(len(p.loss_names), 1), dtype=torch.float32).pin_memory()
                             for p in (self.d_prog, self.g_prog)}
        self.iteration = 0
        self._g_weights_dirty = True
        self.reducers = {}
        if self.world_size > 1 and not SPLIT_GRAPH:
            from ..parallel import GradReducer
            self.reducers = {'d_step': GradReducer(self.d_prog, self.store, ('d',), self.world_size),
                             'g_step': GradReducer(self.g_prog, self.store, ('g', 'c'), self.world_size)}
        self.store.on_load.append(lambda: setattr(self, '_g_weights_dirty', True))

    # ------------------------------------------------------------------ steps
    def _push_lr(self, key, value):
        i = self._lr_pos
        self._lr_pos = (i + 1) % self._lr_ring.numel()
        self._lr_ring[i] = value
        self.lr_dev[key].copy_(self._lr_ring[i:i + 1], non_blocking=True)

    def _allreduce(self, group):
        from ..parallel import allreduce_sum_
        allreduce_sum_(group.grads, self.world_size)

    def _body_a(self, prog, keys, refresh_foreign=True):
        for k in keys:
            if k in self.groups:
                g = self.groups[k]
                _C.call('rcgan_zero', g.grads.data_ptr(), g.numel * 4, _C.stream_ptr())
        prog.run_forward(refresh_foreign)
        prog.run_backward()

    def _body_b(self, prog, keys):
        for k in keys:
            if k in self.groups:
                adam_step(self.groups[k], self.lr_dev[k], 0.0, 0.9, 1e-8, 1.0 / self.world_size)
        prog.run_updates()

    def _run(self, name, fn):
        if not self.use_cuda_graph:
            fn()
            return
        g = self._graphs.get(name)
        if g is None:
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

    def _snapshot(self):
        snap = {n: v.data.clone() for n, v in self.store.vars.items()}
        for k, g in self.groups.items():
            snap['__m_' + k], snap['__v_' + k] = g.m.clone(), g.v.clone()
        return snap

    def _restore(self, snap):
        for n, v in self.store.vars.items():
            v.data.copy_(snap[n])
        for k, g in self.groups.items():
            g.m.copy_(snap['__m_' + k]); g.v.copy_(snap['__v_' + k])

    def _step(self, prog, keys, tag, lrs, refresh_foreign=True):
        if self.rng == 'device':
            # the in-graph random inputs are keyed by the step count, read from device memory by the captured graph
            self._step_count += 1
            i = self._step_count % self._step_ring.numel()
            self._step_ring[i] = self._step_count
            self.rng_step.copy_(self._step_ring[i:i + 1], non_blocking=True)
        for k in keys:
            if k in self.groups:
                self.groups[k].t += 1
                self._push_lr(k, tf_adam_lr(lrs[k], 0.0, 0.9, self.groups[k].t))
        if not refresh_foreign:
            tag += '_keep'            # a second captured graph: the other optimizer's folds / weight packs are not refreshed
        if self.world_size > 1 and SPLIT_GRAPH:
            # (A/B switch RCGAN_DP_SPLIT_GRAPH=1: the round-1 form, one blocking all-reduce per arena between two graph halves)
            self._run(tag + '_a', lambda: self._body_a(prog, keys, refresh_foreign))
            for k in keys:
                if k in self.groups:
                    self._allreduce(self.groups[k])
            self._run(tag + '_b', lambda: self._body_b(prog, keys))
            return
        # one graph per step; with world_size > 1 the gradient buckets are all-reduced from inside the backward sweep
        # (parallel.GradReducer hooks) and joined in front of Adam
        red = self.reducers.get(prog.name)

        def body():
            self._body_a(prog, keys, refresh_foreign)
            if red is not None:
                red.wait()
            self._body_b(prog, keys)
        self._run(tag, body)

    def d_step(self, it=None):
        it = self.iteration if it is None else it
        # the generator's folded filters / weight packs inside the D program only change with a G step (or a load): of the
        # N_CRITIC discriminator steps of an iteration only the first refreshes them
        refresh, self._g_weights_dirty = self._g_weights_dirty, False
        self._step(self.d_prog, ('d',), 'd', {'d': self.FLAGS.lr * lr_decay(it)}, refresh_foreign=refresh)

    def g_step(self, it=None):
        it = self.iteration if it is None else it
        F = self.FLAGS
        clr = F.lr * F.confuse_multiplier * (lr_decay(it) if F.confuse_lr_decay else 1.0)
        self._step(self.g_prog, ('g', 'c'), 'g', {'g': F.lr * lr_decay(it), 'c': clr})
        self._g_weights_dirty = True

    def feed(self, prog, **feeds):
        for name, src in feeds.items():
            if src is None or name not in prog.inputs:      # (rng='device': noise / dequant_noise are not inputs)
                continue
            t = prog.inputs[name]
            src = torch.as_tensor(src)
            want = t.data.dtype
            if src.dtype != want:
                src = src.to(want)
            t.data.copy_(src.reshape(-1), non_blocking=True)

    def sample(self, labels, noise=None):
        """fixed_noise_samples / samples_100 (gan_resnet.py:820-850): Generator(len(labels), labels, noise) in the training
        graph's mode (conditional BN over the sample batch -- the reference's generator has no inference mode).
        Returns float32 [N,32,32,3] NHWC in [-1,1]."""
        labels = np.asarray(labels).reshape(-1)
        N = len(labels)
        if getattr(self, '_s_prog_n', None) != N:
            S.set_store(self.store, self.seed)
            self.s_prog = sp = Program('sampler', self.device, self.act_dtype)
            towers, self.net.towers = self.net.towers, 1
            try:
                with sp:
                    nz = sp.input('noise', [N, Z_DIM])
                    lab = sp.input('labels', [N, 1], I32)
                    img = self.net.Generator(N, lab, noise=CastOp(nz, self.act_dtype).y, reuse=True)
                    self._s_out = CastOp(img, _C.F32).y
            finally:
                self.net.towers = towers
            sp.finalize([])
            self._s_prog_n = N
        if noise is None:
            noise = torch.randn(N, Z_DIM)
        self.feed(self.s_prog, noise=noise, labels=labels)
        self.s_prog.run_forward()
        torch.cuda.current_stream().synchronize()
        return self._s_out.torch().float().cpu().numpy().reshape(N, IMG_SIZE, IMG_SIZE, IMG_DIM)

    def fetch_losses(self):
        for p in (self.d_prog, self.g_prog):
            self._host_losses[p.name].copy_(p.losses, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        out = self.d_prog.loss_dict(self._host_losses['d_step'])
        out.update(self.g_prog.loss_dict(self._host_losses['g_step']))
        return out

    def train_iteration(self, d_feeds, g_feeds=None, fetch=True):
        """One iteration of gan_resnet.py:919-947: [G(+C) step if iteration > 0] then N_CRITIC D steps.
        d_feeds: list of N_CRITIC feed dicts (or one dict reused); g_feeds: feed dict of the G step."""
        if self.iteration > 0:
            if g_feeds:
                self.feed(self.g_prog, **g_feeds)
            self.g_step()
        for i in range(N_CRITIC):
            f = d_feeds[i] if isinstance(d_feeds, (list, tuple)) else d_feeds
            if f:
                self.feed(self.d_prog, **f)
            self.d_step()
        self.iteration += 1
        return self.fetch_losses() if fetch else None
