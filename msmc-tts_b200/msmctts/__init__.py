"""msmctts -- drop-in B200-native backend for the MSMC-VQ-GAN training hot path of hhguo/MSMC-TTS.

Same package name, class names, constructor kwargs, forward signatures and state_dict keys as the
reference's `msmctts.networks` tree, so the reference's train.py / train_dist.py / yaml configs drive it.
"""
