"""CPU oracle for the audio-cut separation-and-feature hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The shipped package
(``audio-cut_b200/``) never imports anything from here and fails loudly when
its CUDA library is missing.

What it restates (reference paths are relative to /root/reference):

* ``planner``   - ``src/audio_cut/utils/gpu_pipeline.py:333-375`` (chunk_schedule)
* ``mdx``       - ``src/audio_cut/separation/backends.py:268-281, 299-406`` (pad /
                  window / trim / stem math) and the external MVSEP-MDX23
                  ``Conv_TDF_net_trim_model.stft/istft`` it calls
* ``unet``      - the TFC-TDF U-Net the opaque ONNX file computes
* ``pipeline``  - ``src/vocal_smart_splitter/core/enhanced_vocal_separator.py:366-373,
                  423-437, 456-458`` (chunk slicing, effective region, uniform average)
                  and ``src/audio_cut/analysis/features_cache.py:122-195, 254-335``
* ``features``  - the librosa (>=0.10) functions those files call

PARITY STATUS: **parity unpinned** for the third-party arithmetic.  The
reference holds no golden vector, known-answer test or fixture for separation,
STFT/iSTFT or any feature series, and ``librosa`` /
``onnxruntime`` / the MVSEP ``inference.py`` are not installed in this image.
What IS pinned against the reference's own code run in this container
(``tests/golden/make_golden.py``): ``chunk_schedule``, the
``MDX23OnnxBackend.infer_chunk`` window/pad/trim/stem arithmetic, the
``_separate_with_pipeline`` stitch, ``ChunkFeatureBuilder`` masking / dedupe,
``_compute_mdd_series`` and ``finalize_cut_points``.
"""
