"""Parameters of the hot path -- the values of /root/reference/constants.py that a1-a22 read
(SURVEY.md section 8a, row a24).  Names follow the reference so callers can star-import it the same way.
"""
import numpy as np

from ._lsf_codebook import LSF_BINS_F32

# quantiser (constants.py:5-8)
init_alpha = -300
beta_boundary = 1

# framing (constants.py:9, :25-27)
sample_rate = 16000
frame_length = 512
overlap_each_side = 32

# module-level switches (constants.py:12-22).  The reference selects these by editing the file; here they are
# only DEFAULTS -- every entry point takes the choice explicitly (CodecConfig.resnet_type, lpc flags).
is_pure_time_domain = True
resnet_type = 'gln'
max_amp_tr = 33.461480140686035 if is_pure_time_domain else 22.307652973859113
mu_law_transform = False
conv_mu = 63.0

selected_ind = [8.0, 16.0, 32.0, 128.0]   # mel resolutions (constants.py:28; mfcc_transform uses the ints)

# LPC (constants.py:63-64, :66-119)
lpc_perceptual_weighting_coeff = 0.92
empha_filter_coeff = -0.68
lpc_order = 16                             # neural_speech_coding_module.py:50
lpc_coeff_lsf_bins = [float(v) for v in LSF_BINS_F32.astype(np.float64)]
