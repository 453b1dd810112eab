"""TEST INFRASTRUCTURE (oracle): ``diffusers.utils`` symbols the reference reads
(/root/reference/distrifuser/modules/pp/attn.py:3). True => no LoRA ``scale`` is forwarded (attn.py:67,120)."""
USE_PEFT_BACKEND = True
