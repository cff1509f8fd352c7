"""DDIMSampler: importable name only.  In the reference fork the DDIM sampler cannot run against this UNet (its
unconditional input lacks "relations", ddim.py:116 vs openaimodel.py:444 -> KeyError), so it is not on the hot path;
callers import the name (txt2img.py:16) but use PLMSSampler."""


class DDIMSampler(object):
    def __init__(self, diffusion, model, schedule="linear", alpha_generator_func=None, set_alpha_scale=None):
        self.diffusion, self.model, self.schedule = diffusion, model, schedule
        self.alpha_generator_func, self.set_alpha_scale = alpha_generator_func, set_alpha_scale

    def sample(self, *args, **kwargs):
        raise NotImplementedError("DDIM sampling is unusable in the reference fork (uncond input has no 'relations'); "
                                  "use ldm.models.diffusion.plms.PLMSSampler")
