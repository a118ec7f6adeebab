"""
Checkpoint compatibility.

``PixelNeRFNet.load_state_dict`` accepts, besides its own ``state_dict()``:

* the fork trainer's checkpoint dict (``{"net_state_dict": ..., "epoch": ...}``,
  train/trainlib/trainer.py:593-610) and ``{"state_dict": ...}`` / ``{"model": ...}`` wrappers;
* ``nn.DataParallel`` / DDP saves (``module.`` prefix on every key);
* upstream pixelNeRF ``pixel_nerf_latest`` files: those have ``encoder.model.*`` only, while the
  fork registers the ResNet stages a second time as ``encoder.layers.<i>.*`` (encoder.py:66-80 of the
  reference), so a strict load of an upstream file fails there with "missing keys".  The aliases are
  filled in from the ``encoder.model.*`` entries that share their storage (they are the same tensors);
* non-persistent camera buffers (``poses``, ``focal``, ``c``, ``image_shape``, ``encoder.latent`` ...)
  that some upstream versions saved: dropped when the model does not expect them.

A width mismatch of ``lin_z`` (upstream: 512-wide multi-scale latent, fork single-scale: 256) is
reported with the conf switch that fixes it instead of torch's generic size-mismatch text.
"""
import re

_WRAPPERS = ("net_state_dict", "model_state_dict", "state_dict", "model", "net")
_TRANSIENT = re.compile(r"^(poses|focal|c|image_shape|num_objs|num_views_per_obj|encoder\.latent(_scaling)?|global_encoder\.latent)$")


def unwrap(obj):
    """checkpoint object -> flat {key: tensor}"""
    for _ in range(3):
        if isinstance(obj, dict):
            hit = [k for k in _WRAPPERS if k in obj and isinstance(obj[k], dict)]
            if hit and not any(hasattr(v, "shape") for v in obj.values()):
                obj = obj[hit[0]]
                continue
        break
    if not isinstance(obj, dict):
        raise TypeError("checkpoint does not contain a state dict (got %s)" % type(obj).__name__)
    return obj


def normalize_state_dict(sd, expected):
    """Maps a foreign state dict onto the keys in ``expected`` (the model's own ``state_dict()``).
    Returns a new dict; never modifies tensors."""
    sd = unwrap(sd)
    if sd and all(k.startswith("module.") for k in sd):
        sd = {k[len("module."):]: v for k, v in sd.items()}
    out = {k: v for k, v in sd.items() if k in expected or not _TRANSIENT.match(k)}
    # aliases: keys of the model that are views of one tensor (encoder.layers.<i>.* and encoder.model.*)
    by_storage = {}
    for k, v in expected.items():
        if hasattr(v, "data_ptr") and v.numel() > 0:
            by_storage.setdefault((v.data_ptr(), tuple(v.shape), v.dtype), []).append(k)
    for group in by_storage.values():
        have = [k for k in group if k in out]
        if have:
            for k in group:
                out.setdefault(k, out[have[0]])
    for k, v in out.items():
        if k in expected and ".lin_z." in k and k.endswith("weight") and tuple(v.shape) != tuple(expected[k].shape):
            raise RuntimeError(
                "%s: checkpoint latent width %d, model latent width %d -- the checkpoint was trained with "
                "encoder.use_multi_scale=%s (512 = stages [64,64,128,256] concatenated; 256 = last stage only, the "
                "fork's default); set model.encoder.use_multi_scale accordingly" %
                (k, v.shape[1], expected[k].shape[1], "true" if v.shape[1] > expected[k].shape[1] else "false"))
    return out
