"""Seeded case lists shared by make_golden.py (generator) and the tests."""

CALC_CASES = [
    # name, seed, k, sigma, kwargs for calc_similarity
    ("inverse_cls", 11, 6, 0.6, dict(use_inverse=True, temperature=0.1, use_cls_token=True, ot_temp=0.05)),
    ("inverse_mean", 12, 6, 0.6, dict(use_inverse=True, temperature=0.1, use_cls_token=False, ot_temp=0.05)),
    ("minus", 13, 6, 0.8, dict(use_minus=True, use_inverse=True, use_cls_token=True)),
    ("soft", 14, 5, 0.6, dict(use_soft=True, use_cls_token=True)),
    ("relu", 15, 5, 0.6, dict(use_cls_token=True)),
    ("uniform", 16, 5, 1.0, dict(use_uniform=True)),
    ("minus_part", 17, 6, 0.6, dict(use_minus=True, ot_part=0.5, use_cls_token=True)),
    ("inverse_part", 18, 4, 0.6, dict(use_inverse=True, temperature=0.1, ot_part=0.9, use_cls_token=True)),
]

LOOP_CASES = [
    # name, n, classes, seed, sigma, trunc_nums, use_rollout, flags
    ("rollout_full", 160, 8, 31, 0.6, [0, 5, 20], True, dict(ot_part=1.0)),
    ("inverse_full", 128, 6, 32, 0.8, [0, 16], False, dict(use_inverse=True, temperature=0.1, use_cls_token=True, ot_part=1.0)),
    ("minus_part", 96, 5, 33, 0.6, [0, 10], False, dict(use_minus=True, ot_part=0.5)),
]
