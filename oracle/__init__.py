"""CPU oracle for the SED hot path (TEST INFRASTRUCTURE ONLY).

Everything under ``oracle/`` restates, on the CPU, what the reference
(ariel415el/SoundEventDetection-Pytorch) computes on its hot path.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it -- and only as the checker or the
reported CPU baseline, never as the product path.  The product
(``soundeventdetection-pytorch_b200``) never imports this package and fails
loudly when its CUDA extension is missing.

Parity status: the reference ships no tests or golden vectors
and its log-mel arithmetic lives in librosa, which is neither vendored nor
installable here.  The log-mel oracle is therefore a restatement of librosa's
published algorithm pinned against two independent implementations
(``torch.stft`` and ``torchaudio.functional.melscale_fbanks``) -- see
``tests/test_oracle_logmel.py``.  The CNN oracle is pinned against the verbatim
reference ``nn.Module``s imported from /root/reference (fixtures under
``tests/golden/``, generator ``tests/golden/make_golden.py``).
"""
