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


def test_round2_headline_line_reports_the_tensor_roof_and_the_sub_records():
    """Round 2: the conv kernels are judged against the TENSOR pipe, with the issued fraction, ncu's
    tensor-pipe activity and the measured DRAM traffic beside it; the default line carries bounded sub-records for the other
    BASELINE.json configurations and for the reference's shipped switches ('gln', the_strides 4)."""
    d = _line('r02_bench_default.json')
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
              'dtype', 'data', 'config', 'e2e', 'gpu_launches', 'clocks', 'roofline', 'cpu_baseline', 'sub_records'):
        assert k in d, k
    r = d['roofline']
    assert r['bound'] == 'tensor' and r['unit'] == 'TFLOP/s' and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9
    assert abs(r['issued_frac'] - 3 * r['frac']) < 1e-9 and r['mma_per_product'] == 3          # three MMAs per product (fp16 hi/lo)
    assert 0 < r['tensor_pipe_active_ncu'] <= 100 and r['traffic'] > 0
    assert r['traffic_detail']['traffic_vs_algorithmic'] > 1.0                                  # measured DRAM bytes vs real-channel payload
    ws = r['whole_step']
    assert ws['algorithmic_bytes_per_frame'] == 4624 and ws['traffic_vs_algorithmic'] > 100     # layer-by-layer: the honest number
    assert 0 < r['hbm']['frac_of_measured'] < 1
    e = d['e2e']
    assert e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] > 0 and 0 < e['value'] <= d['value'] * 1.02
    s = d['sub_records']
    for k in ('codec1_b128', 'cq_scaled', 'cq2_gln', 'cq2_stride4', 'cq2_gln_stride4', 'train', 'train_gln', 'corpus_1h_per_gpu'):
        assert k in s and 'error' not in s[k], k
    assert s['codec1_b128']['value'] >= 6000 and s['codec1_b128']['cpu_baseline']['kind'] == 'port'
    assert s['codec1_b128']['launches_per_call'] < s['codec1_b128']['unprepared']['launches_per_call']
    for k in ('cq2_gln', 'cq2_stride4', 'cq2_gln_stride4'):
        assert s[k]['engine'].startswith('plane engine') and s[k]['value'] > 5000, k
    assert s['train']['collectives_per_step'] in (0, 1) and s['train']['ms_per_step'] > 0
    ref = _line('r02_bench_reference.json')
    assert ref['impl'] == 'reference' and ref['metric'] == d['metric'] and ref['unit'] == d['unit']
    assert ref['e2e'] == {'value': ref['value'], 'unit': ref['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_round2_final_lines_state_how_the_weights_were_prepared():
    """The headline is measured with the weights prepared once per weights (nsc_prepare), says so in `config`, and reports the
    unprepared step beside it; the line of the final configuration (4,144-frame passes, dependent launches) carries the same keys and
    is not slower than the plain-launch line beyond run-to-run spread; the traffic of the roofline record is per launch of THAT run."""
    d, f = _line('r02_bench_default.json'), _line('r02_bench_default_final.json')
    for line in (d, f):
        assert 'nsc_prepare' in line['config']['weights']
        u = line['unprepared']
        assert u['ms_per_step'] > 0 and abs(u['ms_per_step'] / line['ms_per_step'] - 1.0) < 0.05
        r = line['roofline']
        assert abs(r['traffic'] / r['hbm']['algorithmic_bytes_per_launch'] - r['traffic_detail']['traffic_vs_algorithmic']) < 0.05
        assert 0 < line['e2e']['value'] <= line['value'] * 1.02
        for k in ('codec1_b128', 'cq_scaled', 'cq2_gln', 'cq2_stride4', 'cq2_gln_stride4', 'train', 'train_gln', 'corpus_1h_per_gpu'):
            assert 'error' not in line['sub_records'][k], k
    assert f['value'] > 0.97 * d['value'] and f['sub_records']['corpus_1h_per_gpu']['value'] > 8000
    assert 'cpu_baseline' in d and d['cpu_baseline']['kind'] == 'port'
