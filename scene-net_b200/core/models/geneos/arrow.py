"""Cone-on-cylinder GENEOs — mirror of core/models/geneos/arrow.py:30-252 over the CUDA synthesis.

cone_kernel (v1): gaussian ring exp(-(d^2-r^2)^2/(2 sig^2)), cone slices use
sig_h = cone_radius*sin(cone_inc*pi/(2+h));  arrow (v2): sigma*exp(-d^4/(2(rad+1e-8)^2)), cone
slices use rad_h = cone_radius*h*tan(clamp(cone_inc,0,0.499)*pi).  int(apex) cylinder planes
sit at the END of the z axis, the cone slices at the beginning (arrow.py:241-250).
"""
import torch

from .GENEO_kernel_torch import GENEO_kernel_torch


class cone_kernel(GENEO_kernel_torch):
    kind_name = "cone_kernel"
    abi_params = ("apex", "cone_inc", "cone_radius", "radius", "sigma")

    def __init__(self, name, kernel_size, plot=False, **kwargs):
        if kwargs.get('radius') is None:
            raise KeyError("Provide a radius for the cylinder in the kernel.")
        if kwargs.get('apex') is None:
            raise KeyError("Provide a height for the cone.")
        if kwargs.get('cone_inc') is None:
            raise KeyError("Provide an inclination for the cone.")
        self.radius = kwargs['radius']
        self.apex = kwargs['apex']
        self.cone_inc = kwargs['cone_inc']
        self.cone_radius = kwargs['cone_radius'] if kwargs.get('cone_radius') is not None \
            else torch.tensor(float(kernel_size[1] - 1))
        self.sigma = kwargs['sigma'] if kwargs.get('sigma') is not None else torch.tensor(1.0)
        hc = int(float(self.apex))
        if hc < 0 or hc > int(kernel_size[0]):
            raise ValueError(f"int(apex) = {hc} must lie in [0, kernel_size[0] = {kernel_size[0]}]")
        if plot:
            print("--- Cone Kernel ---")
            print(f"radius = {float(self.radius):.4f}; apex = {float(self.apex):.4f}; "
                  f"cone_radius = {float(self.cone_radius):.4f}; cone_inc = {float(self.cone_inc):.4f}")
        super().__init__(name, kernel_size, plot)

    def mandatory_parameters():
        return ['radius', 'apex', 'cone_radius', 'cone_inc']

    def geneo_parameters():
        return cone_kernel.mandatory_parameters() + ['sigma']

    def geneo_random_config(name='GENEO_rand'):
        cfg = GENEO_kernel_torch.geneo_random_config()
        k = cfg['kernel_size']
        # same draws, same order as arrow.py:122-128
        cfg['geneo_params'] = {
            'radius': torch.randint(1, k[1], (1,))[0] / 2,
            'apex': torch.randint(int(k[0] / 2), k[0] - 1, (1,))[0],
            'cone_radius': torch.randint(1, k[1], (1,))[0] / 2,
            'cone_inc': torch.rand(1, )[0],
            'sigma': torch.randint(5, 10, (1,))[0] / 5,
        }
        cfg['name'] = 'cone'
        cfg['non_trainable'] = ['apex']
        return cfg

    def geneo_smart_config(name="Smart_Cylinder"):
        return {'name': name, 'kernel_size': (9, 6, 6), 'plot': False, 'non_trainable': [],
                'geneo_params': {'radius': torch.tensor(1.0), 'apex': torch.tensor(3.0), 'cone_radius': torch.tensor(2.0),
                                 'cone_inc': torch.tensor(0.1), 'sigma': torch.tensor(2.0)}}


class arrow(cone_kernel):
    kind_name = "arrow"

    def __init__(self, name, kernel_size, plot=False, **kwargs):
        super().__init__(name, kernel_size, plot, **kwargs)
