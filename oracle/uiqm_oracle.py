"""numpy restatement of getUIQM (uqim_utils.py:176-190) — TEST INFRASTRUCTURE (oracle/__init__.py).

Vectorised where the reference loops, same arithmetic order where it matters; keeps the
reference's lambda_b = 0.144 (uqim_utils.py:107)."""
import math

import numpy as np
from scipy import ndimage


def _mu_a(x, a_l=0.1, a_r=0.1):
    """uqim_utils.py:10-28 (note the reference starts at index T_a_L + 1)."""
    x = np.sort(x)
    k = len(x)
    t_l = math.ceil(a_l * k)
    t_r = math.floor(a_r * k)
    return float(np.sum(x[t_l + 1:k - t_r].astype(np.float64)) / (k - t_l - t_r))


def _uicm(img):
    """uqim_utils.py:36-48."""
    r, g, b = (img[:, :, i].flatten() for i in range(3))
    rg = r - g
    yb = (r + g) / 2 - b
    m_rg, m_yb = _mu_a(rg), _mu_a(yb)
    s_rg = float(np.mean((rg.astype(np.float64) - m_rg) ** 2))
    s_yb = float(np.mean((yb.astype(np.float64) - m_yb) ** 2))
    return -0.0268 * math.sqrt(m_rg ** 2 + m_yb ** 2) + 0.1586 * math.sqrt(s_rg + s_yb)


def _sobel(x):
    """uqim_utils.py:50-55."""
    mag = np.hypot(ndimage.sobel(x, 0), ndimage.sobel(x, 1))
    mag *= 255.0 / np.max(mag)
    return mag


def _eme(x, ws):
    """uqim_utils.py:57-82."""
    k1, k2 = x.shape[1] // ws, x.shape[0] // ws
    blocks = x[:ws * k2, :ws * k1].reshape(k2, ws, k1, ws)
    mx = blocks.max(axis=(1, 3))
    mn = blocks.min(axis=(1, 3))
    ok = (mn != 0.0) & (mx != 0.0)
    val = 0.0
    for l in range(k1):            # same accumulation order as the reference loops
        for k in range(k2):
            if ok[k, l]:
                val += math.log(mx[k, l] / mn[k, l])
    return 2.0 / (k1 * k2) * val


def _uism(img):
    """uqim_utils.py:84-108."""
    emes = [_eme(_sobel(img[:, :, c]) * img[:, :, c], 10) for c in range(3)]
    return 0.299 * emes[0] + 0.587 * emes[1] + 0.144 * emes[2]


def _uiconm(img, ws):
    """uqim_utils.py:141-174."""
    k1, k2 = img.shape[1] // ws, img.shape[0] // ws
    blocks = img[:ws * k2, :ws * k1].reshape(k2, ws, k1, ws, img.shape[2])
    mx = blocks.max(axis=(1, 3, 4))
    mn = blocks.min(axis=(1, 3, 4))
    val = 0.0
    for l in range(k1):
        for k in range(k2):
            top, bot = mx[k, l] - mn[k, l], mx[k, l] + mn[k, l]
            if math.isnan(top) or math.isnan(bot) or bot == 0.0 or top == 0.0:
                continue
            val += (top / bot) * math.log(top / bot)
    return -1.0 / (k1 * k2) * val


def get_uiqm(img):
    """uqim_utils.py:176-190: returns (uiqm, uicm, uism, uiconm) for an HxWx3 0..255 array."""
    x = img.astype(np.float32)
    uicm, uism, uiconm = _uicm(x), _uism(x), _uiconm(x, 10)
    return 0.0282 * uicm + 0.2953 * uism + 3.5753 * uiconm, uicm, uism, uiconm
