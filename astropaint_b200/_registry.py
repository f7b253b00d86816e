"""Marks Python templates that have a CUDA device-function twin (csrc/profiles.cuh).

`Painter` looks for `template._ap_device`; templates without it go through the flat work-list
path (device geometry + scatter, template evaluated by the user's Python on the host), or the
table path when they are wrapped in `@interpolate`.
"""

# ids of include/astropaint_b200.h :: enum ap_profile
AP_CONSTANT, AP_LINEAR, AP_SOLID_SPHERE, AP_GAUSSIAN = 0, 1, 2, 3
AP_NFW_RHO2D, AP_NFW_TAU2D, AP_NFW_KSZ, AP_NFW_DEFLECTION, AP_NFW_BG = 4, 5, 6, 7, 8
AP_B16_TAU2D, AP_B16_RHOGAS2D, AP_NFW_RHO2D_LOS = 9, 10, 11


def device_profile(profile_id, args, kwonly=None):
    """args: template argument names (catalog columns or spray kwargs) in device order;
    kwonly: {name: default} keyword-only arguments appended after them."""
    def mark(fun):
        fun._ap_device = {"id": profile_id, "args": list(args), "kwonly": dict(kwonly or {})}
        return fun
    return mark


def device_table(kind, args, n_samples, sampling_method="linspace", scale=None):
    """A template that is `@interpolate`-tabulated with node values computed on device.
    scale(columns dict of torch tensors) -> per-halo multiplier tensor, or None."""
    def mark(fun):
        fun._ap_device_table = {"kind": kind, "args": list(args), "n_samples": n_samples,
                                "sampling_method": sampling_method, "scale": scale}
        return fun
    return mark
