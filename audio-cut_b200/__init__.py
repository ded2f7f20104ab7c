"""B200-native separation-and-feature hot path of BDMstudio/audio-cut (drop-in).

Import name: ``audio_cut_b200`` (the directory is ``audio-cut_b200/``; the sibling
``audio_cut_b200/`` shim makes it importable).
"""
__version__ = "0.1.0"
