# the default bit widths of the RPXMB multishuffle (sums to 64); placeholder module, see sxs/__init__.py
default_shuffle_widths = (8, 8, 4, 4, 4, 2) + (1,) * 34
