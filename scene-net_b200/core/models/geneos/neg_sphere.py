"""Negative-sphere GENEOs — mirror of core/models/geneos/neg_sphere.py:29-199 over the CUDA synthesis.

neg_sphere_kernel (v1): exp(-(d^2-r^2)^2/(2 sigma^2)) - mean - neg_factor;
negSpherev2: K = -neg_factor*sigma*exp(-d^4/(2(r+1e-8)^2)), K -= (sum K + neg_factor)/T.
The reference's index layout (a genuine scramble for non-cubic kernels) is reproduced.
"""
import torch

from .GENEO_kernel_torch import GENEO_kernel_torch


class neg_sphere_kernel(GENEO_kernel_torch):
    kind_name = "neg_sphere_kernel"
    abi_params = ("neg_factor", "radius", "sigma")

    def __init__(self, name, kernel_size, plot=False, **kwargs):
        if kwargs.get('radius') is None:
            raise KeyError("Provide a radius for the sphere.")
        if kwargs.get('neg_factor') is None:
            raise KeyError("Provide a negative factor for each sphere weight.")
        self.radius = kwargs['radius']
        self.neg_factor = kwargs['neg_factor']
        self.sigma = kwargs['sigma'] if kwargs.get('sigma') is not None else torch.tensor(1.0)
        if plot:
            print("--- Neg. Sphere Kernel ---")
            print(f"radius = {float(self.radius):.4f}; neg_factor = {float(self.neg_factor):.4f}")
        super().__init__(name, kernel_size)

    def mandatory_parameters():
        return ['radius', 'neg_factor']

    def geneo_parameters():
        return neg_sphere_kernel.mandatory_parameters() + ['sigma']

    def geneo_random_config(name='GENEO_rand'):
        cfg = GENEO_kernel_torch.geneo_random_config()
        # same draws, same order as neg_sphere.py:92-96
        cfg['geneo_params'] = {
            'radius': torch.randint(1, cfg['kernel_size'][1], (1,))[0],
            'neg_factor': torch.randint(1, 10, (1,))[0] / 10,
            'sigma': torch.randint(5, 10, (1,))[0] / 10,
        }
        cfg['non_trainable'] = []
        cfg['name'] = 'neg'
        return cfg

    def geneo_smart_config(name="Smart_Neg_Sphere"):
        return {'name': name, 'kernel_size': (9, 6, 6), 'plot': False, 'non_trainable': [],
                'geneo_params': {'radius': torch.tensor(3.0), 'sigma': torch.tensor(2.0), 'neg_factor': torch.tensor(0.5)}}


class negSpherev2(neg_sphere_kernel):
    kind_name = "negSpherev2"

    def __init__(self, name, kernel_size, **kwargs):
        super().__init__(name, kernel_size, **kwargs)
