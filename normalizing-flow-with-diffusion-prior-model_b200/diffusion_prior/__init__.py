"""diffusion_prior — only the step either side of the Glow inverse in the NFDPM sampling path:
the latent formaters.  The DDPM UNet, its trainer and the Gaussian-diffusion loop of the reference package stay the
reference's own files; importing them from here raises AttributeError by design."""
from .latent_formaters import BaseFormater, IdentityFormater, CatFormater, get_formater

__all__ = ["BaseFormater", "IdentityFormater", "CatFormater", "get_formater"]
