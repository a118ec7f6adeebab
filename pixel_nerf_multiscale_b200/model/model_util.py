"""Factories (mirror of src/model/model_util.py)."""
from .encoder import ImageEncoder, SpatialEncoder
from .resnetfc import ResnetFC


def make_mlp(conf, d_in, d_latent=0, allow_empty=False, **kwargs):
    mlp_type = conf.get_string("type", "mlp")
    if mlp_type == "resnet":
        return ResnetFC.from_conf(conf, d_in, d_latent=d_latent, **kwargs)
    if mlp_type == "empty" and allow_empty:
        return None
    # 'mlp' (ImplicitNet) is unreachable in the reference as well (NameError, model_util.py:8)
    raise NotImplementedError("Unsupported MLP type %s" % mlp_type)


def make_encoder(conf, **kwargs):
    enc_type = conf.get_string("type", "spatial")
    if enc_type == "spatial":
        return SpatialEncoder.from_conf(conf, **kwargs)
    if enc_type == "global":
        return ImageEncoder.from_conf(conf, **kwargs)
    raise NotImplementedError("Unsupported encoder type")
