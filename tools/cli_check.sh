#!/bin/bash
# End-to-end check of the reference-compatible command line on a GPU box: toy training with a TensorFlow-bundle
# checkpoint written and read back, validation, and the demo path on a synthetic two-tone wav.
set -e
cd "$(dirname "$0")/.."
rm -rf gpurun_out/cli && mkdir -p gpurun_out/cli
python main.py -n cli -m train -ne 2 -bs 2 --no-save-on-epoch -o gpurun_out/cli/ckpt 2>&1 | tail -4
ls gpurun_out/cli
python main.py -n cli -m valid -bs 2 -i gpurun_out/cli/ckpt 2>&1 | tail -2
python - <<'PY'
import numpy as np, scipy.io.wavfile
t = np.arange(16000) / 8000.
w = 3000. * np.sin(2 * np.pi * 440. * t) + 2000. * np.sin(2 * np.pi * 1200. * t) * (0.5 + 0.5 * np.cos(2 * np.pi * 3. * t))
scipy.io.wavfile.write('gpurun_out/cli/mix.wav', 8000, w.astype(np.int16))
PY
(cd gpurun_out/cli && python ../../main.py -n cli -m demo -i ckpt -if mix.wav 2>&1 | tail -3 && ls)
python - <<'PY'
import numpy as np, scipy.io.wavfile
a = scipy.io.wavfile.read('gpurun_out/cli/demo_0.wav')[1]; b = scipy.io.wavfile.read('gpurun_out/cli/demo_1.wav')[1]
print('demo outputs', a.shape, b.shape, 'finite', bool(np.isfinite(a).all() and np.isfinite(b).all()), 'rms %.1f %.1f' % (np.sqrt((a**2).mean()), np.sqrt((b**2).mean())))
PY
