"""Duck-typed stand-ins for the gpytorch / botorch objects that appear in the reference's
signatures (likelihood, covar_module, priors, constraints, datasets, posteriors).

gpytorch and botorch are not installable in this image, and the arithmetic they would do is
exactly what the CUDA kernels replace; what the reference API needs from them is only their
*parameter containers*: raw parameters, Interval constraints, priors, initial values
(scamlgp/model.py:25-105).  These classes hold that state, convert it to the C ABI's
`scaml_hyper_spec` + raw parameter row, and keep the attribute names the reference uses
(`.raw_lengthscale`, `.lengthscale`, `.outputscale`, `.noise`, `.state_dict()` ...).
"""
from __future__ import annotations

import copy
import math
from typing import Dict, Iterator, Optional, Tuple

import torch

from ._capi import (KERNEL_MATERN12, KERNEL_MATERN32, KERNEL_MATERN52, KERNEL_RBF, PRIOR_GAMMA, PRIOR_LOGNORMAL,
                    PRIOR_NONE, HyperSpec)

DT = torch.float64


class ModelFittingError(RuntimeError):
    """botorch.exceptions.errors.ModelFittingError stand-in (raised by optimize_marginal_likelihood)."""


# ---- constraints -------------------------------------------------------------------------- #
class Interval:
    """gpytorch.constraints.Interval: value = lower + (upper - lower) * sigmoid(raw)."""

    def __init__(self, lower_bound, upper_bound, initial_value=None):
        self.lower_bound = float(lower_bound)
        self.upper_bound = float(upper_bound)
        self.initial_value = None if initial_value is None else float(initial_value)

    def transform(self, raw: torch.Tensor) -> torch.Tensor:
        return self.lower_bound + (self.upper_bound - self.lower_bound) * torch.sigmoid(raw)

    def inverse_transform(self, value: torch.Tensor) -> torch.Tensor:
        u = (torch.as_tensor(value, dtype=DT) - self.lower_bound) / (self.upper_bound - self.lower_bound)
        return torch.log(u) - torch.log1p(-u)


class GreaterThan:
    """gpytorch.constraints.GreaterThan(lower, transform=None): the raw value IS the value; the bound is
    enforced by the optimiser (reference scamlgp/model.py:333-337)."""

    def __init__(self, lower_bound, transform=None, initial_value=None):
        if transform is not None:
            raise NotImplementedError("only transform=None (box bound) is supported")
        self.lower_bound = float(lower_bound)
        self.initial_value = initial_value

    def transform(self, raw):
        return raw

    def inverse_transform(self, value):
        return torch.as_tensor(value, dtype=DT)


# ---- priors ------------------------------------------------------------------------------- #
class GammaPrior:
    def __init__(self, concentration: float, rate: float):
        self.concentration, self.rate = float(concentration), float(rate)

    def spec(self) -> Tuple[int, float, float]:
        return (PRIOR_GAMMA, self.concentration, self.rate)

    def log_prob(self, x):
        a, b = self.concentration, self.rate
        return a * math.log(b) + (a - 1.0) * torch.log(x) - b * x - math.lgamma(a)

    def sample(self, shape, generator=None) -> torch.Tensor:
        # Marsaglia-Tsang through torch's own sampler; generator-aware via standard_gamma
        conc = torch.full(tuple(shape), self.concentration, dtype=DT)
        return torch._standard_gamma(conc, generator=generator) / self.rate if generator is not None else \
            torch._standard_gamma(conc) / self.rate


class LogNormalPrior:
    def __init__(self, loc: float, scale: float):
        self.loc, self.scale = float(loc), float(scale)

    def spec(self) -> Tuple[int, float, float]:
        return (PRIOR_LOGNORMAL, self.loc, self.scale)

    def log_prob(self, x):
        lx = torch.log(x)
        return -lx - math.log(self.scale) - 0.5 * math.log(2 * math.pi) - (lx - self.loc) ** 2 / (2 * self.scale ** 2)

    def sample(self, shape, generator=None) -> torch.Tensor:
        z = torch.randn(tuple(shape), dtype=DT, generator=generator)
        return torch.exp(self.loc + self.scale * z)


def _prior_spec(prior) -> Tuple[int, float, float]:
    return (PRIOR_NONE, 0.0, 0.0) if prior is None else prior.spec()


# ---- kernels / likelihood ----------------------------------------------------------------- #
class _ParamModule:
    """Minimal nn.Module-like container: named raw parameters, state_dict, deepcopy."""

    _param_names: Tuple[str, ...] = ()

    def state_dict(self, prefix: str = "") -> Dict[str, torch.Tensor]:
        return {prefix + n: getattr(self, n).clone() for n in self._param_names}

    def load_state_dict(self, sd: Dict[str, torch.Tensor], prefix: str = "") -> None:
        for n in self._param_names:
            if prefix + n in sd:
                setattr(self, n, sd[prefix + n].clone().to(DT))


class RBFKernel(_ParamModule):
    kernel_id = KERNEL_RBF
    _param_names = ("raw_lengthscale",)

    def __init__(self, ard_num_dims: Optional[int] = None, batch_shape=torch.Size(), lengthscale_prior=None,
                 lengthscale_constraint: Optional[Interval] = None):
        if len(tuple(batch_shape)) != 0:
            raise NotImplementedError("batch_shape must be () (the reference optimizer always passes torch.Size())")
        self.ard_num_dims = ard_num_dims
        self.lengthscale_prior = lengthscale_prior
        self.raw_lengthscale_constraint = lengthscale_constraint or Interval(1e-4, 1e2, 0.5)
        c = self.raw_lengthscale_constraint
        init = c.initial_value if c.initial_value is not None else 0.5 * (c.lower_bound + c.upper_bound)
        d = 1 if ard_num_dims is None else ard_num_dims
        self.raw_lengthscale = c.inverse_transform(torch.full((1, d), init, dtype=DT))

    @property
    def lengthscale(self) -> torch.Tensor:
        return self.raw_lengthscale_constraint.transform(self.raw_lengthscale)

    @lengthscale.setter
    def lengthscale(self, value):
        v = torch.as_tensor(value, dtype=DT).reshape(1, -1).expand_as(self.raw_lengthscale)
        self.raw_lengthscale = self.raw_lengthscale_constraint.inverse_transform(v).clone()


class MaternKernel(RBFKernel):
    def __init__(self, nu: float = 2.5, **kwargs):
        super().__init__(**kwargs)
        ids = {0.5: KERNEL_MATERN12, 1.5: KERNEL_MATERN32, 2.5: KERNEL_MATERN52}
        if nu not in ids:
            raise ValueError("nu must be one of 0.5, 1.5, 2.5")
        self.nu = nu
        self.kernel_id = ids[nu]


class ScaleKernel(_ParamModule):
    _param_names = ("raw_outputscale",)

    def __init__(self, base_kernel: RBFKernel, batch_shape=torch.Size(), outputscale_prior=None,
                 outputscale_constraint: Optional[Interval] = None):
        self.base_kernel = base_kernel
        self.outputscale_prior = outputscale_prior
        self.raw_outputscale_constraint = outputscale_constraint or Interval(1e-4, 1e2, 1.0)
        c = self.raw_outputscale_constraint
        init = c.initial_value if c.initial_value is not None else 1.0
        self.raw_outputscale = c.inverse_transform(torch.tensor(init, dtype=DT))

    @property
    def outputscale(self) -> torch.Tensor:
        return self.raw_outputscale_constraint.transform(self.raw_outputscale)

    @outputscale.setter
    def outputscale(self, value):
        self.raw_outputscale = self.raw_outputscale_constraint.inverse_transform(torch.as_tensor(value, dtype=DT)).clone()

    def state_dict(self, prefix: str = ""):
        sd = super().state_dict(prefix)
        sd.update(self.base_kernel.state_dict(prefix + "base_kernel."))
        return sd

    def load_state_dict(self, sd, prefix: str = ""):
        super().load_state_dict(sd, prefix)
        self.base_kernel.load_state_dict(sd, prefix + "base_kernel.")


class GaussianLikelihood(_ParamModule):
    _param_names = ("raw_noise",)

    def __init__(self, noise_prior=None, noise_constraint: Optional[Interval] = None, batch_shape=torch.Size()):
        self.noise_prior = noise_prior
        self.raw_noise_constraint = noise_constraint or Interval(1e-8, 1e-2, 1e-3)
        c = self.raw_noise_constraint
        init = c.initial_value if c.initial_value is not None else 1e-3
        self.raw_noise = c.inverse_transform(torch.full((1,), init, dtype=DT))

    @property
    def noise(self) -> torch.Tensor:
        return self.raw_noise_constraint.transform(self.raw_noise)

    @noise.setter
    def noise(self, value):
        self.raw_noise = self.raw_noise_constraint.inverse_transform(torch.as_tensor(value, dtype=DT).reshape(1)).clone()


# ---- conversion to / from the C ABI view ---------------------------------------------------- #
def _unwrap(covar_module):
    if isinstance(covar_module, ScaleKernel):
        return covar_module, covar_module.base_kernel
    raise TypeError("covar_module must be ScaleKernel(RBFKernel | MaternKernel) -- the kernel family the reference "
                    "constructs (scamlgp/model.py:44-70, 87-105)")


def hyper_spec_of(likelihood: GaussianLikelihood, covar_module: ScaleKernel) -> HyperSpec:
    sk, bk = _unwrap(covar_module)
    lc, oc, nc = bk.raw_lengthscale_constraint, sk.raw_outputscale_constraint, likelihood.raw_noise_constraint
    return HyperSpec(
        kernel=bk.kernel_id,
        ls_bounds=(lc.lower_bound, lc.upper_bound), os_bounds=(oc.lower_bound, oc.upper_bound),
        noise_bounds=(nc.lower_bound, nc.upper_bound),
        ls_prior=_prior_spec(bk.lengthscale_prior), os_prior=_prior_spec(sk.outputscale_prior),
        noise_prior=_prior_spec(likelihood.noise_prior),
        ls_init=float(bk.lengthscale.flatten()[0]), os_init=float(sk.outputscale), noise_init=float(likelihood.noise[0]),
    )


def theta_raw_of(likelihood: GaussianLikelihood, covar_module: ScaleKernel, d: int) -> torch.Tensor:
    """[raw lengthscales (d), raw outputscale, raw noise] -- the C ABI's parameter row."""
    sk, bk = _unwrap(covar_module)
    ls = bk.raw_lengthscale.reshape(-1)
    if ls.numel() == 1 and d > 1:
        ls = ls.expand(d)
    if ls.numel() != d:
        raise ValueError(f"kernel has {ls.numel()} lengthscales but the data has {d} input dimensions")
    return torch.cat([ls, sk.raw_outputscale.reshape(1), likelihood.raw_noise.reshape(1)]).to(DT)


def set_theta_raw(likelihood: GaussianLikelihood, covar_module: ScaleKernel, theta_raw: torch.Tensor) -> None:
    sk, bk = _unwrap(covar_module)
    t = theta_raw.detach().to("cpu", DT)
    d = t.numel() - 2
    bk.raw_lengthscale = t[:d].reshape(1, d).clone()
    bk.ard_num_dims = d
    sk.raw_outputscale = t[d].clone()
    likelihood.raw_noise = t[d + 1].reshape(1).clone()


def named_priors_of(likelihood, covar_module) -> Iterator[Tuple[str, object, Interval, int]]:
    """(name, prior, constraint, #values) in the order of the parameter row."""
    sk, bk = _unwrap(covar_module)
    yield "covar_module.base_kernel.lengthscale_prior", bk.lengthscale_prior, bk.raw_lengthscale_constraint
    yield "covar_module.outputscale_prior", sk.outputscale_prior, sk.raw_outputscale_constraint
    yield "likelihood.noise_covar.noise_prior", likelihood.noise_prior, likelihood.raw_noise_constraint


# ---- data containers ------------------------------------------------------------------------ #
class _CallableTensor(torch.Tensor):
    """botorch 0.7.3's SupervisedDataset exposes X / Y both as attributes with `.shape` and as callables
    (`task_data.X()`, scamlgp/model.py:180-181 vs `data.X.shape`, utils.py:117-118)."""

    # attribute reads (`.shape`) and arithmetic on the container's tensors take the plain-Tensor path: without this every
    # `task_data.X.shape` of validate_meta_data went through the __torch_function__ protocol (75 ms at 4096 tasks)
    __torch_function__ = torch._C._disabled_torch_function_impl

    def __call__(self):
        return self.as_subclass(torch.Tensor)


class SupervisedDataset:
    def __init__(self, X: torch.Tensor, Y: torch.Tensor):
        self.X = torch.as_tensor(X).as_subclass(_CallableTensor)
        self.Y = torch.as_tensor(Y).as_subclass(_CallableTensor)


class Standardize:
    """botorch Standardize(m=1) state: means / stdvs (unbiased; < 1e-8 or NaN -> 1)."""

    def __init__(self, m: int = 1, batch_shape=torch.Size()):
        self.means = torch.zeros(1, 1, dtype=DT)
        self.stdvs = torch.ones(1, 1, dtype=DT)
        self._is_trained = False
        self.training = True

    def fit(self, Y: torch.Tensor) -> "Standardize":
        Y = Y.reshape(-1).to(DT)
        self.means = Y.mean().reshape(1, 1)
        s = Y.std(unbiased=True) if Y.numel() > 1 else torch.tensor(float("nan"), dtype=DT)
        self.stdvs = (s if bool(s >= 1e-8) else torch.ones((), dtype=DT)).reshape(1, 1)
        self._is_trained = True
        return self

    def __call__(self, Y: torch.Tensor, Yvar=None):
        if self.training:
            self.fit(Y)
        return (Y - self.means.to(Y.device)) / self.stdvs.to(Y.device), Yvar

    def untransform(self, Y: torch.Tensor, Yvar=None):
        return self.means.to(Y.device) + self.stdvs.to(Y.device) * Y, Yvar

    def eval(self):
        self.training = False
        return self

    def train(self, mode: bool = True):
        self.training = mode
        return self


class MultivariateNormal:
    def __init__(self, mean: torch.Tensor, covariance_matrix: torch.Tensor):
        self.mean = mean
        self.covariance_matrix = covariance_matrix
        self.lazy_covariance_matrix = covariance_matrix

    @property
    def variance(self):
        return self.covariance_matrix.diagonal(dim1=-2, dim2=-1)


class Posterior:
    """botorch GPyTorchPosterior view: `.mean`, `.variance` with a trailing output dimension, `.mvn`."""

    def __init__(self, mean: torch.Tensor, variance: torch.Tensor, covariance: Optional[torch.Tensor] = None):
        self._mean, self._var, self._cov = mean, variance, covariance

    @property
    def mean(self):
        return self._mean.unsqueeze(-1)

    @property
    def variance(self):
        return self._var.unsqueeze(-1)

    @property
    def mvn(self):
        cov = self._cov if self._cov is not None else torch.diag_embed(self._var)
        return MultivariateNormal(self._mean, cov)


def clone_module(m):
    return copy.deepcopy(m)
