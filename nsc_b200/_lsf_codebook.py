"""The 256-entry LSF codebook initialiser of the reference (constants.py:66-119, `lpc_coeff_lsf_bins`).

It is DATA the hot path is parameterised by: 16 concatenated linear ramps
of 16 values each -- NOT monotone over the whole table.  Stored as the little-endian float32 image
(base64) of the reference's decimal literals so the values are bit-identical to what
``tf.Variable(lpc_coeff_lsf_bins, dtype=tf.float32)`` holds (cmrl.py:781, nscm.py:997).
tests/test_oracle_pins.py (test_lsf_codebook_matches_reference_constants) re-derives it from /root/reference/constants.py when
that tree is present.
"""
import base64

import numpy as np

_B64 = (
    "kInwPAjxhz2sv9M9J8cPPnmuNT7MlVs+j76APjeykz7gpaY+iZm5PjKNzD7bgN8+hHTyPha0Aj/rLQw/v6cVP4fyuj2NwBA+1wdE"
    "PiBPdz41S5U+2u6uPn+SyD4jNuI+yNn7Pre+Cj+JkBc/W2IkPy40MT8ABj4/09dKP6WpVz+pO2U+DpWQPkiMrj6Cg8w+u3rqPvs4"
    "BD+XNBM/NDAiP9ErMT9uJ0A/CyNPP6geXj9EGm0/4RV8P7+IhT+NBo0/BrKEPl5apz62Aso+DqvsPrOpBz/f/Rg/DFIqPzimOz9k"
    "+kw/kE5eP7yibz90e4A/iiWJP6DPkT+2eZo/zCOjP1sMzT5fLvA+MagJPzM5Gz80yiw/Nls+PzfsTz85fWE/Ow5zP55Pgj8fGIs/"
    "oOCTPyGpnD+icaU/IjquP6MCtz+T9Ao/ySIdP/5QLz80f0E/aa1TP57bZT/UCXg/BRyFPx8zjj86Spc/VWGgP3B4qT+Kj7I/paa7"
    "P8C9xD/b1M0/dZk1P4K3Rz+O1Vk/m/NrP6cRfj/aF4g/4CaRP+Y1mj/sRKM/8lOsP/litT//cb4/BYHHPwuQ0D8Rn9k/GK7iPzZE"
    "dz8uloM/QYqLP1V+kz9ocps/e2ajP45aqz+hTrM/tEK7P8g2wz/bKss/7h7TPwET2z8UB+M/J/vqPzvv8j+mkJE/H6iZP5i/oT8R"
    "16k/iu6xPwMGuj98HcI/9jTKP29M0j/oY9o/YXviP9qS6j9TqvI/zMH6P6JsAUBfeAVANNmjP7lmrD899LQ/woG9P0YPxj/LnM4/"
    "TyrXP9S33z9YReg/3dLwP2Fg+T/z9gBAtT0FQHiECUA6yw1A/BESQGfxuj8iocM/3VDMP5cA1T9SsN0/DWDmP8cP7z+Cv/c/njcA"
    "QHyPBEBZ5whANj8NQBSXEUDx7hVAzkYaQKyeHkBmUdA/wDfYPxke4D9zBOg/zervPyfR9z+At/8/7c4DQBrCB0BHtQtAdKgPQKCb"
    "E0DNjhdA+oEbQCd1H0BUaCNAcurmP8qG7j8iI/Y/er/9P+mtAkAVfAZAQkoKQG4YDkCa5hFAxrQVQPKCGUAeUR1ASh8hQHbtJECi"
    "uyhAz4ksQKm1AEBqXgRAKgcIQOuvC0CrWA9AbAETQCyqFkDtUhpArvsdQG6kIUAvTSVA7/UoQLCeLEBwRzBAMfAzQPGYN0AJTBBA"
    "BoITQAS4FkAC7hlA/yMdQP1ZIED6jyNA+MUmQPb7KUDzMS1A8WcwQO6dM0Ds0zZA6gk6QOc/PUDldUBAzWojQEfBJUDBFyhAO24q"
    "QLXELEAvGy9AqXExQCPIM0CdHjZAFnU4QJDLOkAKIj1AhHg/QP7OQUB4JURA8ntGQA=="
)

LSF_BINS_F32 = np.frombuffer(base64.b64decode(_B64), dtype="<f4").copy()
assert LSF_BINS_F32.shape == (256,)
