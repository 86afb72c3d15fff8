"""Grid NeRF field - host-side mirror of the reference's ``nerf/network_grid.py`` (:13-181).

Tiled multi-resolution grid encoder -> MLP(32 -> 64 -> 64 -> 4) -> ``trunc_exp`` density (plus a
Gaussian blob at the origin) and ``sigmoid`` albedo; frequency-encoded 2-layer background MLP;
finite-difference normals + Lambertian shading.  Module / parameter names match the reference so
its checkpoints load (``encoder.embeddings``, ``sigma_net.net.{0,1,2}.{weight,bias}``,
``bg_net.net.{0,1}.*``).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import field as _field
from .activation import trunc_exp
from .encoding import get_encoder
from .renderer import NeRFRenderer, safe_normalize


class MLP(nn.Module):
    def __init__(self, dim_in, dim_out, dim_hidden, num_layers, bias=True):
        super().__init__()
        self.dim_in, self.dim_out, self.dim_hidden, self.num_layers = dim_in, dim_out, dim_hidden, num_layers
        self.net = nn.ModuleList([
            nn.Linear(dim_in if l == 0 else dim_hidden, dim_out if l == num_layers - 1 else dim_hidden, bias=bias)
            for l in range(num_layers)])

    def forward(self, x):
        for l, layer in enumerate(self.net):
            x = layer(x)
            if l != self.num_layers - 1:
                x = F.relu(x, inplace=True)
        return x


class NeRFNetwork(NeRFRenderer):
    def __init__(self, opt, num_layers=3, hidden_dim=64, num_layers_bg=2, hidden_dim_bg=64):
        super().__init__(opt)
        self.num_layers = num_layers
        self.hidden_dim = hidden_dim

        self.encoder, self.in_dim = get_encoder('tiledgrid', input_dim=3, log2_hashmap_size=16,
                                                desired_resolution=2048 * self.bound)
        self.sigma_net = MLP(self.in_dim, 4, hidden_dim, num_layers, bias=True)

        if self.bg_radius > 0:
            self.num_layers_bg = num_layers_bg
            self.hidden_dim_bg = hidden_dim_bg
            self.encoder_bg, self.in_dim_bg = get_encoder('frequency', input_dim=3)
            self.bg_net = MLP(self.in_dim_bg, 3, hidden_dim_bg, num_layers_bg, bias=True)
        else:
            self.bg_net = None

    def gaussian(self, x):
        # density blob at the scene centre: 5 * exp(-|x|^2 / (2 * 0.2^2))  (network_grid.py:66-74)
        d = (x ** 2).sum(-1)
        return 5 * torch.exp(-d / (2 * 0.2 ** 2))

    fused = True  # use the fused tcgen05 field kernels when the shapes / autocast state allow it

    def common_forward(self, x):
        # x: [N, 3] in [-bound, bound] -> sigma [N] (fp32), albedo [N, 3]
        if self.fused and _field.can_fuse(x, self.encoder, self.sigma_net):
            return _field.fused_field(x, self.encoder, self.sigma_net, self.bound)
        h = self.encoder(x, bound=self.bound)
        h = self.sigma_net(h)
        sigma = trunc_exp(h[..., 0] + self.gaussian(x))
        albedo = torch.sigmoid(h[..., 1:])
        return sigma, albedo

    def finite_difference_normal(self, x, epsilon=1e-2):
        # central differences of the density along each axis (network_grid.py:90-105)
        grads = []
        for axis in range(3):
            off = torch.zeros(1, 3, device=x.device)
            off[0, axis] = epsilon
            pos, _ = self.common_forward((x + off).clamp(-self.bound, self.bound))
            neg, _ = self.common_forward((x - off).clamp(-self.bound, self.bound))
            grads.append(0.5 * (pos - neg) / epsilon)
        return -torch.stack(grads, dim=-1)

    def _stencil_fusable(self, x):
        return self.fused and _field.can_fuse(x.detach() if x.requires_grad else x, self.encoder, self.sigma_net) and not x.requires_grad

    def normal(self, x):
        if self._stencil_fusable(x):
            # the six shifted points of every sample as six consecutive rows of ONE fused-field batch (csrc/shading.cu)
            from . import step_ops
            sigma6, _ = _field.fused_field(step_ops.stencil_points(x, 1e-2, self.bound, with_centre=False), self.encoder,
                                           self.sigma_net, self.bound)
            return step_ops.stencil_normal(sigma6)
        normal = safe_normalize(self.finite_difference_normal(x))
        normal[torch.isnan(normal)] = 0
        return normal

    def forward(self, x, d, l=None, ratio=1, shading='albedo'):
        # x: [N, 3]; d: [N, 3] view dirs; l: [3] light dir; ratio: ambient ratio
        if shading == 'albedo':
            sigma, color = self.common_forward(x)
            normal = None
        elif self._stencil_fusable(x):
            # 7-point stencil: centre + six shifted points per sample in one fused-field launch, then one kernel for
            # normal + lambertian / textureless / normal colouring (csrc/shading.cu)
            from . import step_ops
            M = x.shape[0]
            sigma_all, rgb_all = _field.fused_field(step_ops.stencil_points(x, 1e-2, self.bound, with_centre=True), self.encoder,
                                                    self.sigma_net, self.bound)
            normal, color = step_ops.shade(sigma_all, rgb_all, l, ratio, shading)
            sigma = sigma_all.view(M, 7)[:, 0]
        else:
            sigma, albedo = self.common_forward(x)
            normal = self.normal(x)
            lambertian = ratio + (1 - ratio) * (normal @ l).clamp(min=0)  # [N]
            if shading == 'textureless':
                color = lambertian.unsqueeze(-1).repeat(1, 3)
            elif shading == 'normal':
                color = (normal + 1) / 2
            else:  # 'lambertian'
                color = albedo * lambertian.unsqueeze(-1)
        return sigma, color, normal

    def density(self, x):
        sigma, albedo = self.common_forward(x)
        return {'sigma': sigma, 'albedo': albedo}

    def background(self, d):
        if self.fused:
            from . import step_ops
            if step_ops.can_fuse_background(d, self.encoder_bg, self.bg_net):
                return step_ops.background_net(d, self.bg_net)
        h = self.encoder_bg(d)  # [N, C]
        h = self.bg_net(h)
        return torch.sigmoid(h)

    def get_params(self, lr):
        params = [
            {'params': self.encoder.parameters(), 'lr': lr * 10},
            {'params': self.sigma_net.parameters(), 'lr': lr},
        ]
        if self.bg_radius > 0:
            params.append({'params': self.encoder_bg.parameters(), 'lr': lr * 10})
            params.append({'params': self.bg_net.parameters(), 'lr': lr})
        return params
