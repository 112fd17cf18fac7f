python tools/attn_bwd_cta_life.py cross
python tools/attn_bwd_cta_life.py dec
