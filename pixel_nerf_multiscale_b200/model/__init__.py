from .models import PixelNeRFNet


def make_model(conf, *args, **kwargs):
    """conf['model'] -> PixelNeRFNet (mirror of src/model/__init__.py:7-14)."""
    model_type = conf.get_string("type", "pixelnerf")
    if model_type == "pixelnerf":
        return PixelNeRFNet(conf, *args, **kwargs)
    raise NotImplementedError("Unsupported model type", model_type)
