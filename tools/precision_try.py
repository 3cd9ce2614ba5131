"""Measures the encoder (floating code) and decoder (same codes) error of the CUDA codec against the float64 oracle."""
import sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from oracle import ref_codec
from util import ar_frames, rel_err
from nsc_b200 import codec
import test_gpu_parity as tp
ocfg=ref_codec.OracleCodecCfg(); cfg=codec.CodecConfig(resnet_type='bottleneck', precision=sys.argv[1] if len(sys.argv)>1 else 'tc_f16x3')
for seed in (3,4,7):
    oc=ref_codec.OracleCodec(ocfg, seed=seed)
    gc=codec.NeuralCodec(cfg, torch.from_numpy(codec.pack_params_numpy(cfg, oc.conv_params, oc.alpha, oc.bins)).cuda())
    x=ar_frames(16,512,seed=60+seed,std=0.3)
    r=gc.computational_graph_end2end_quan_on(torch.from_numpy(x).cuda(), True, 1.0)
    o64=oc.forward(torch.from_numpy(x).double()[:,:,None], True, 1.0)
    e_code=rel_err(r['floating_code'].cpu().numpy(), o64['floating_code'].numpy()[:,:,0])
    code_g=r['code'].cpu().numpy()
    oc.ps._cursor=tp._enc_layers(oc)
    out_o=oc.decoder(torch.from_numpy(code_g).double()[:,:,None])[:,:,0].numpy()
    e_dec=rel_err(r['out'].cpu().numpy(), out_o)
    print(f'seed {seed}: floating code {e_code:.2e}  decoder on identical codes {e_dec:.2e}')
