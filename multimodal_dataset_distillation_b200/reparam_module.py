"""ReparamModule: run an nn.Module with an external flat fp32 parameter vector.

Mirror of the reference's interface (reparam_module.py:9-159): constructor ``ReparamModule(module)``,
``forward(*inputs, flat_param=None, buffers=None, **kw)``, attributes ``flat_param`` (nn.Parameter),
``param_numel``, ``_param_infos``, ``_param_numels``, ``_param_shapes``, ``_shared_param_infos``,
``_buffer_infos``.  Flattening order = ``named_modules()`` x ``named_parameters(recurse=False)``
(reparam_module.py:28-51), shared parameters are stored once; buffers are not reparametrised.

The flat vector this class produces for ``ProjectionHead`` is exactly the layout the CUDA kernels take
(``ops.head_numel`` / include/vldd_b200.h), so ``txt_student_net.flat_param`` can be handed to
``ops.unrolled_match`` unchanged.  ``trace()`` of the reference is unused by the hot path and not provided.
"""
from __future__ import annotations

from contextlib import contextmanager

import torch
import torch.nn as nn


def _resolve(root: nn.Module, path: str) -> nn.Module:
    mod = root
    if path:
        for part in path.split("."):
            mod = getattr(mod, part)
    return mod


class ReparamModule(nn.Module):
    def __init__(self, module: nn.Module):
        super().__init__()
        self.module = module
        infos, shared, tensors, seen = [], [], [], {}
        for mod_name, mod in self.named_modules():
            for p_name, p in mod.named_parameters(recurse=False):
                if p is None:
                    continue
                if p in seen:
                    shared.append((mod_name, p_name) + seen[p])
                else:
                    seen[p] = (mod_name, p_name)
                    infos.append((mod_name, p_name))
                    tensors.append(p.detach())
        if len({t.dtype for t in tensors}) > 1:
            raise AssertionError("expects all parameters in module to have same dtype")
        self._param_infos = tuple(infos)
        self._shared_param_infos = tuple(shared)
        self._param_numels = tuple(t.numel() for t in tensors)
        self._param_shapes = tuple(t.size() for t in tensors)
        flat = torch.cat([t.reshape(-1) for t in tensors], 0) if tensors else torch.zeros(0)
        self.register_parameter("flat_param", nn.Parameter(flat))
        self.param_numel = flat.numel()
        # the named parameters become plain attributes that alias views of a flat vector
        for mod_name, p_name in self._param_infos:
            delattr(_resolve(self, mod_name), p_name)
        for mod_name, p_name, _, _ in self._shared_param_infos:
            delattr(_resolve(self, mod_name), p_name)
        self._unflatten_param(self.flat_param)
        self._buffer_infos = tuple((mn, n, b) for mn, m in self.named_modules()
                                   for n, b in m.named_buffers(recurse=False) if b is not None)

    # ---- parameter views ------------------------------------------------------------------------
    def _views(self, flat_param: torch.Tensor):
        return [chunk.view(shape) for chunk, shape in zip(flat_param.split(self._param_numels), self._param_shapes)]

    def _install(self, views) -> None:
        for (mod_name, p_name), v in zip(self._param_infos, views):
            setattr(_resolve(self, mod_name), p_name, v)
        for mod_name, p_name, src_mod, src_name in self._shared_param_infos:
            setattr(_resolve(self, mod_name), p_name, getattr(_resolve(self, src_mod), src_name))

    def _unflatten_param(self, flat_param: torch.Tensor) -> None:
        self._install(self._views(flat_param))

    def clear_views(self) -> None:
        for mod_name, p_name in self._param_infos:
            setattr(_resolve(self, mod_name), p_name, None)

    @contextmanager
    def unflattened_param(self, flat_param: torch.Tensor):
        saved = [getattr(_resolve(self, mn), n) for mn, n in self._param_infos]
        self._install(self._views(flat_param))
        try:
            yield
        finally:
            self._install(saved)

    @contextmanager
    def replaced_buffers(self, buffers):
        for (mn, n, _), new in zip(self._buffer_infos, buffers):
            setattr(_resolve(self, mn), n, new)
        try:
            yield
        finally:
            for mn, n, old in self._buffer_infos:
                setattr(_resolve(self, mn), n, old)

    # ---- forward --------------------------------------------------------------------------------
    def _forward_with_param(self, flat_param, *inputs, **kwinputs):
        with self.unflattened_param(flat_param):
            return self.module(*inputs, **kwinputs)

    def _forward_with_param_and_buffers(self, flat_param, buffers, *inputs, **kwinputs):
        with self.unflattened_param(flat_param), self.replaced_buffers(buffers):
            return self.module(*inputs, **kwinputs)

    def forward(self, *inputs, flat_param=None, buffers=None, **kwinputs):
        if flat_param is None:
            flat_param = self.flat_param
        else:
            flat_param = torch.squeeze(flat_param)      # DataParallel hands each replica a [1, P] slice (distill.py:516-517)
        if buffers is None:
            return self._forward_with_param(flat_param, *inputs, **kwinputs)
        return self._forward_with_param_and_buffers(flat_param, tuple(buffers), *inputs, **kwinputs)
