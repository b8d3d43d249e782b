"""No-op matplotlib stand-in (TEST INFRASTRUCTURE ONLY): the reference only draws diagnostic PNGs."""
