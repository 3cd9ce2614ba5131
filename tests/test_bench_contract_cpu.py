"""The committed bench lines under profiles/ carry every key the bench contract asks for (no GPU: reads the JSON evidence)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, 'profiles', name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_headline_line_has_the_contract_keys():
    d = _line('r01_bench_plane_default.json')
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
              'dtype', 'data', 'config', 'e2e', 'gpu_launches', 'clocks', 'roofline', 'cpu_baseline'):
        assert k in d, k
    assert d['higher_is_better'] is True and d['scaling'] == 'weak' and d['vs_baseline'] is None and d['warmup'] >= 3
    assert 'workload' in d['config'] and 'model' not in d['config']
    e = d['e2e']
    assert e['unit'] == d['unit'] and e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] > 0 and 0 < e['value'] <= d['value'] * 1.02
    assert d['gpu_launches'] > 0
    r = d['roofline']
    assert r['bound'] in ('hbm', 'tensor') and r['unit'] in ('GB/s', 'TFLOP/s')
    assert abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9 and 0 < r['frac'] < 1
    assert isinstance(r['traffic'], (int, float)) and r['traffic'] > 0           # DRAM bytes of one launch from the ncu capture
    c = d['cpu_baseline']
    assert c['kind'] in ('port', 'reference') and c['cores'] >= 1 and c['value'] > 0 and c['unit'] == d['unit'] and c['sample']
    assert set(d['clocks']) >= {'sm_mhz', 'sm_max_mhz', 'reasons'}
    assert not set(d['clocks']['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}


def test_reference_arm_line_matches_the_headline_arm():
    d, r = _line('r01_bench_plane_default.json'), _line('r01_bench_reference.json')
    assert r['impl'] == 'reference'
    for k in ('metric', 'unit', 'higher_is_better'):
        assert r[k] == d[k], k
    assert r['e2e'] == {'value': r['value'], 'unit': r['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert r['cpu_baseline']['value'] == r['value'] and r['cpu_baseline']['kind'] == 'port'
