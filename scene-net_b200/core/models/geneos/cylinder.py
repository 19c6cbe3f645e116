"""Cylinder GENEOs — mirror of core/models/geneos/cylinder.py:30-176 over the CUDA synthesis.

cylinder_kernel: exp(-(d^2 - r^2)^2 / (2 sigma^2));  cylinderv2: sigma * exp(-d^4 / (2 (r+1e-8)^2));
both zero-summed over the (x,y) plane and tiled along z.
"""
import torch

from .GENEO_kernel_torch import GENEO_kernel_torch, _as_param


class cylinder_kernel(GENEO_kernel_torch):
    kind_name = "cylinder_kernel"
    abi_params = ("radius", "sigma")

    def __init__(self, name, kernel_size, plot=False, **kwargs):
        if kwargs.get('radius') is None:
            raise KeyError("Provide a radius for the cylinder in the kernel.")
        self.radius = kwargs['radius']
        self.sigma = kwargs['sigma'] if kwargs.get('sigma') is not None else torch.tensor(1.0)
        if plot:
            print("--- Cylinder Kernel ---")
            print(f"radius = {float(self.radius):.4f}; sigma = {float(self.sigma):.4f}")
        super().__init__(name, kernel_size)

    def mandatory_parameters():
        return ['radius']

    def geneo_parameters():
        return cylinder_kernel.mandatory_parameters() + ['sigma']

    def geneo_random_config(name='GENEO_rand'):
        cfg = GENEO_kernel_torch.geneo_random_config()
        # same draws, same order as cylinder.py:115-118 (keeps the RNG stream of a seeded run)
        cfg['geneo_params'] = {
            'radius': torch.randint(1, cfg['kernel_size'][1], (1,))[0] / 2,
            'sigma': torch.randint(5, 10, (1,))[0] / 5,
        }
        cfg['name'] = 'cylinder'
        return cfg

    def geneo_smart_config(name="Smart_Cylinder"):
        return {'name': name, 'kernel_size': (9, 6, 6), 'plot': False, 'non_trainable': [],
                'geneo_params': {'radius': torch.tensor(1.0), 'sigma': torch.tensor(2.0)}}


class cylinderv2(cylinder_kernel):
    kind_name = "cylinderv2"

    def __init__(self, name, kernel_size, plot=False, **kwargs):
        super().__init__(name, kernel_size, plot, **kwargs)
