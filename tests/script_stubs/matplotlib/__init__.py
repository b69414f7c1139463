"""Stand-in for matplotlib: ddpm_3d_ldm/show_model.py only draws mid-slice grids
(show_model.py:104-153)."""


def use(_backend):
    return None
