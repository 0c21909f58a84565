"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the diffusion super-resolution hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the checker
or the timed CPU baseline -- never as a fallback for the CUDA path.

Contents
--------
``schedule.py``        float64 numpy restatement of the beta schedules and the 12 DDPM buffers.
``nets.py``            functional fp32 torch-CPU restatement of the ResDiff / SRDiff UNets, the RRDB encoder
                       and the SimpleCNN prior (state_dict in, tensors out; no nn.Module from the reference).
``process.py``         the reverse (p_sample) chain with injected noise and the q_sample + loss training step.
``ref_shims.py``       the three analysis shims that let the *real* reference be imported in a container that
                       has ``/root/reference`` (pytorch_wavelets Haar stand-in, matplotlib stub, CPU ``.cuda()``).
``make_golden.py``     runs the real reference (through the shims) on seeded inputs/weights and writes the small
                       fixtures committed under ``tests/golden/``.

Parity pinning: the reference ships no tests or golden vectors, so the oracle is pinned
against outputs of the reference itself, run in the build container by ``make_golden.py``; the fixtures and the
generating script are committed.  The single un-verifiable assumption is the sign/order convention of
``pytorch_wavelets.DWTForward`` (package not installable here; see ``ref_shims.py``).
"""
