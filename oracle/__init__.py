"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's per-chunk codec.

Nothing under mtscomp_b200/ imports this package.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs
(`cpu_baseline`, `--impl reference`) may use it, and only as the checker / the timed CPU baseline.
"""
