"""CPU oracle for the captioning / GPT-2 training step — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package, and only as the checker or the timed CPU baseline.  The product path (gpt2-vision-language_b200/)
never imports it and has no CPU fallback.
"""
