"""Run ONE sub-record of bench.py (debugging aid): python tools/subrec_probe.py <name>"""
import sys, os, json, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from nsc_b200 import _lib
args = argparse.Namespace(precision='tc_f16x3', train_batch=128, utterances=360, utt_seconds=10.0, no_cpu_baseline=True, steps=2, warmup=1)
lib = _lib.load()
name = sys.argv[1]
dev = 'cuda:0'
fn = {
    'codec1': lambda: bench.measure_codec1(args, 1, 0, dev, lib, cpu=False),
    'sweep': lambda: bench.measure_cq_sweep(args, 1, 0, dev, lib),
    'gln': lambda: bench.measure_variant(args, 1, 0, dev, lib, 'gln', (2,)),
    'stride4': lambda: bench.measure_variant(args, 1, 0, dev, lib, 'bottleneck', (2, 2)),
    'gln_stride4': lambda: bench.measure_variant(args, 1, 0, dev, lib, 'gln', (2, 2)),
    'train': lambda: bench.measure_train(args, 1, 0, dev, lib, 5, 2, breakdown=False),
    'corpus': lambda: bench.measure_corpus(args, 1, 0, dev, lib, 1, 1, n_utt=360),
}[name]
r = fn()
print(name, 'ok', json.dumps(r)[:200])
