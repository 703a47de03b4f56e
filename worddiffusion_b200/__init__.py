"""worddiffusion_b200 -- B200-native (sm_100a) denoising hot path of WordDiffusion.

Drop-in modules (same constructor / forward / state_dict as the reference):
    worddiffusion_b200.unet.UNetModel, worddiffusion_b200.unetPhosc.UNetModelPhosc,
    worddiffusion_b200.unetPhosc2.UNetModelPhosc
Sampling loops: worddiffusion_b200.diffusion.Diffusion (DDPM, DDIM, the reference's reduced-call generator)
PHOSC labels on the device: worddiffusion_b200.phosc.phosc_labels
Training step: worddiffusion_b200.training.FusedTrainStep
C ABI: include/wd_b200.h, implemented by worddiffusion_b200/_lib/libwd_b200.so (build: python -m worddiffusion_b200.build)
"""
__version__ = "0.1.0"
