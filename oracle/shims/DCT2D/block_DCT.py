"""Shadow of DCT2D.block_DCT (imported at src/2D-DCT.py:17-20) -> oracle."""
from oracle import vcf_oracle as _o

def analyze_image(img, block_y_side, block_x_side):
    return _o.analyze_image_loop(img, block_y_side, block_x_side)

def synthesize_image(img, block_y_side, block_x_side):
    return _o.synthesize_image_loop(img, block_y_side, block_x_side)

def get_subbands(img, block_y_side, block_x_side):
    return _o.get_subbands(img, block_y_side, block_x_side)

def get_blocks(img, block_y_side, block_x_side):
    return _o.get_blocks(img, block_y_side, block_x_side)

analyze_block = _o.analyze_block
synthesize_block = _o.synthesize_block
