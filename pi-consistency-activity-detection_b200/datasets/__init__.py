"""Stand-ins for the reference's ``datasets/`` modules so its scripts import and run where the UCF101-24 / JHMDB
videos and pickle splits are absent (the real loaders need skvideo / OpenCV decoding of files that
are not in this environment and are OUT of the hot-path scope).

Every class keeps the reference's constructor signature and sample format (``ucf_dataloader.py:189`` dict keys,
tensor shapes and value ranges) and serves deterministic SYNTHETIC clips.  A regular package, so it wins over both
the reference's namespace directory and the unrelated HuggingFace ``datasets`` in site-packages when this
directory's parent is on ``PYTHONPATH`` (``b200caps.launch`` arranges that)."""
