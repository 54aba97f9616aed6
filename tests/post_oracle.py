"""numpy restatement of CRender's export-time post chain — TEST INFRASTRUCTURE (the checker for SURVEY.md
§8f row N4). Follows src/render/post/post_processor.cpp:110-311 and assets/app/shaders/post_process.comp,
blur.comp, including the quirks that change pixels (8x8 dispatch with integer division, the (i % w, i / h)
addressing of _brightness, the 0.7162 gray weight). float32 throughout, like the shaders."""
import numpy as np

W = np.asarray([0.19744746769063704, 0.1746973469158936, 0.12099884565428047, 0.06560233156931679, 0.027839605612666265, 0.009246250740395456,
                0.002403157286908872, 0.00048872837522002], np.float32)
f = np.float32


def _brightness(img, threshold):
    h, w = img.shape[:2]
    out = img.copy()
    for i in range(w * h):  # post_processor.cpp:302-308
        x, y = i % w, i // h
        if y >= h:
            continue
        at = out[y, x, :3]
        if (at[0] * f(0.2126) + at[1] * f(0.7152)) + at[2] * f(0.0722) < f(threshold):
            out[y, x] = (0, 0, 0, 1)
    return out


def _blur_pass(src, horizontal):
    h, w = src.shape[:2]
    gw, gh = 8 * (w // 8), 8 * (h // 8)  # glDispatchCompute(ceil(w/8), ceil(h/8)) with integer division
    dst = np.zeros_like(src)
    pad = np.zeros((h + 14, w + 14, 3), np.float32)  # texelFetch outside the image reads 0
    pad[7 : 7 + h, 7 : 7 + w] = src[..., :3]
    acc = pad[7 : 7 + h, 7 : 7 + w] * W[0]
    for i in range(1, 8):
        if horizontal:
            acc = acc + pad[7 : 7 + h, 7 + i : 7 + i + w] * W[i]
            acc = acc + pad[7 : 7 + h, 7 - i : 7 - i + w] * W[i]
        else:
            acc = acc + pad[7 + i : 7 + i + h, 7 : 7 + w] * W[i]
            acc = acc + pad[7 - i : 7 - i + h, 7 : 7 + w] * W[i]
    dst[:gh, :gw, :3] = acc[:gh, :gw]
    dst[:gh, :gw, 3] = 1
    return dst


def _um(x):
    A, B, C, D, E, F = f(0.15), f(0.50), f(0.10), f(0.20), f(0.02), f(0.30)
    return ((x * (A * x + C * B) + D * E) / (x * (A * x + B) + D * F)) - E / F


def process(img, use_bloom=False, bloom_threshold=0.7, bloom_strength=1.0, use_gray_scale=False, use_tonemapping=False, tonemapping_type=0,
            tonemapping_exposure=1.0, gamma_correction=2.2):
    img = np.asarray(img, np.float32)
    if not (use_bloom or use_gray_scale or use_tonemapping):
        return img.copy()
    h, w = img.shape[:2]
    gw, gh = 8 * (w // 8), 8 * (h // 8)
    c = img[..., :3].copy()
    if use_bloom:
        b = _brightness(img, bloom_threshold)
        for i in range(10):
            b = _blur_pass(b, i % 2 == 0)
        c = c + b[..., :3] * f(bloom_strength)
    if use_gray_scale:
        gs = (f(0.2126) * c[..., 0] + f(0.7162) * c[..., 1]) + f(0.0722) * c[..., 2]
        c = np.stack([gs, gs, gs], -1)
    if use_tonemapping:
        e, ig = f(tonemapping_exposure), f(1.0) / f(gamma_correction)
        with np.errstate(invalid="ignore", divide="ignore"):
            if tonemapping_type == 0:
                c = np.power(c * e, ig)
            elif tonemapping_type == 1:
                c = c * e
                c = np.power(c / (c + f(1.0)), ig)
            elif tonemapping_type == 2:
                x = np.maximum(f(0.0), c * e - f(0.004))
                c = (x * (f(6.2) * x + f(0.5))) / (x * (f(6.2) * x + f(1.7)) + f(0.06))
            elif tonemapping_type == 3:
                ws = f(1.0) / _um(f(11.2))
                c = np.power(_um(f(2.0) * (c * e)) * ws, ig)
    out = np.zeros_like(img)
    out[:gh, :gw, :3] = c[:gh, :gw].astype(np.float32)
    out[:gh, :gw, 3] = 1
    return out
