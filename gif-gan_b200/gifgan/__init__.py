"""gifgan -- B200-native conv-GAN training step behind gif-gan's Python entry points
(/root/reference/models/recurrent_z/{ops,model,z_model_lib,main,z_model,model_sampler}.py)."""
__version__ = "0.1.0"
