#!/usr/bin/env python
"""
Thin driver with the reference's command line (main.py:551-740): same flags, same config layering
(defaults -> -c JSON -> CLI overrides -> digest), same modes `train | valid | test | demo`, driving
`danet_tensorflow_b200.Model`, i.e. the registered Encoder / Estimator / Separator plugins on the
sm_100a kernels.  This is a harness for the hot path, not a port of the reference's training loop
(TensorBoard summaries, NaN rollback and the interactive/debug modes are out of scope).
"""
import argparse
import sys

import numpy as np
import torch

import danet_tensorflow_b200 as D
from danet_tensorflow_b200 import hparams


def batches(dataset, subset, B, C, train_len=None):
    """main.py:414-426: [B*C, T, F] spectra -> [B, C, T, F] complex, optional random crop"""
    for (sig,) in dataset.epoch(subset, B * C, shuffle=(subset == 'train')):
        sig = torch.as_tensor(sig)
        if not torch.is_complex(sig):
            sig = torch.complex(sig.float(), torch.zeros_like(sig, dtype=torch.float32))
        sig = sig.to('cuda', torch.complex64).reshape(B, C, -1, hparams.FEATURE_SIZE)
        if train_len is not None and sig.shape[2] > train_len:
            t0 = np.random.randint(0, sig.shape[2] - train_len + 1)
            sig = sig[:, :, t0:t0 + train_len]
        yield sig.contiguous()


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('-n', '--name', default='UnnamedExperiment')
    ap.add_argument('-m', '--mode', default='train', help='"train", "valid", "test" or "demo"')
    ap.add_argument('-i', '--input-pfile')
    ap.add_argument('-o', '--output-pfile')
    ap.add_argument('-c', '--hparams-file')
    ap.add_argument('-ne', '--num-epoch', type=int, default=10)
    ap.add_argument('--no-save-on-epoch', action='store_true')
    ap.add_argument('--no-valid-on-epoch', action='store_true')
    ap.add_argument('-if', '--input-file', help='input WAV file for "demo" mode')
    ap.add_argument('-ds', '--dataset')
    ap.add_argument('-lr', '--learn-rate')
    ap.add_argument('-tl', '--train-length')
    ap.add_argument('-bs', '--batch-size')
    args = ap.parse_args(argv)

    if args.hparams_file is not None:
        hparams.load_json(args.hparams_file)
    if args.learn_rate is not None:
        hparams.LR = float(args.learn_rate)
        assert hparams.LR >= 0.
    if args.train_length is not None:
        hparams.MAX_TRAIN_LEN = int(args.train_length)
        assert hparams.MAX_TRAIN_LEN >= 2
    if args.dataset is not None:
        hparams.DATASET_TYPE = args.dataset
    if args.batch_size is not None:
        hparams.BATCH_SIZE = int(args.batch_size)
        assert hparams.BATCH_SIZE > 0
    if args.mode == 'demo':
        hparams.BATCH_SIZE = 1                                   # main.py:623-627
    hparams.digest()

    sys.stdout.write('Preparing dataset "%s" ... ' % hparams.DATASET_TYPE)
    dataset = hparams.get_dataset()()
    dataset.install_and_load()
    sys.stdout.write('done\n')

    model = D.Model(args.name).build()
    model.reset()
    if args.input_pfile is not None:
        model.load_params_file(args.input_pfile)          # *.pt or a TensorFlow checkpoint prefix (main.py:201-206)
    print('%d parameters' % model.parameter_count())
    B, C = hparams.BATCH_SIZE, hparams.MAX_N_SIGNAL

    def sweep(subset):
        loss, snr, n = 0., 0., 0
        for src in batches(dataset, subset, B, C):
            out = model.train_forward(src)
            loss += float(out['valid_loss'])
            snr += float(out['valid_snr'])
            n += 1
        print('%s: loss=%.4e SNR=%.3f dB' % (subset, loss / max(n, 1), snr / max(n, 1)))

    if args.mode == 'train':
        if hparams.LR_DECAY_TYPE not in (None, 'adaptive', 'fixed'):
            raise ValueError('Unknown LR_DECAY_TYPE "%s"' % hparams.LR_DECAY_TYPE)          # main.py:449-451
        best_loss, stale_epochs = float('inf'), 0
        for epoch in range(1, args.num_epoch + 1):
            loss, snr, n = 0., 0., 0
            for src in batches(dataset, 'train', B, C, hparams.MAX_TRAIN_LEN):
                out = model.train_step(src)
                loss += float(out['loss'])
                snr += float(out['snr'])
                n += 1
                sys.stdout.write('.')
                sys.stdout.flush()
            print('\nEpoch %d/%d: loss=%.4e SNR=%.3f dB LR=%g' % (epoch, args.num_epoch, loss / n, snr / n, hparams.LR))
            # learning-rate schedule (main.py:438-459): 'adaptive' counts epochs without a new best mean loss, 'fixed'
            # counts every epoch; after NUM_EPOCH_PER_LR_DECAY of them the rate is multiplied by LR_DECAY
            if hparams.LR_DECAY_TYPE == 'adaptive':
                if loss / n < best_loss:
                    best_loss, stale_epochs = loss / n, 0
                else:
                    stale_epochs += 1
            elif hparams.LR_DECAY_TYPE == 'fixed':
                stale_epochs += 1
            if hparams.LR_DECAY_TYPE is not None and stale_epochs == hparams.NUM_EPOCH_PER_LR_DECAY:
                stale_epochs = 0
                old_lr = model.get_learn_rate()
                model.set_learn_rate(old_lr * hparams.LR_DECAY)
                print('[LR %f -> %f]' % (old_lr, model.get_learn_rate()))
            if not args.no_save_on_epoch:
                model.save_params('saves_%s_e%d.pt' % (args.name, epoch))
            if not args.no_valid_on_epoch:
                sweep('valid')
        if args.output_pfile is not None:
            model.save_params(args.output_pfile)
    elif args.mode in ('valid', 'test'):
        sweep(args.mode)
    elif args.mode == 'demo':
        if args.input_file is None:
            raise ValueError('demo mode needs -if <wav file>')
        import scipy.io.wavfile
        import scipy.signal
        rate, data = scipy.io.wavfile.read(args.input_file)       # app/utils.py:111-116
        data = data.astype(np.float32)
        if rate != hparams.SMPRATE:
            data = scipy.signal.resample(data, int(np.ceil(len(data) * hparams.SMPRATE / rate))).astype(np.float32)
        wavs = model.separate(torch.from_numpy(data[None]).cuda())[0].cpu().numpy()
        for c, w in enumerate(wavs):
            scipy.io.wavfile.write('demo_%d.wav' % c, hparams.SMPRATE, w.astype(np.float32))
            print('wrote demo_%d.wav' % c)
    else:
        raise ValueError('Unknown mode "%s"' % args.mode)       # main.py:739


if __name__ == '__main__':
    main()
