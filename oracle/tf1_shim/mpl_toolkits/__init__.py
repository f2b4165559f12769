"""TEST INFRASTRUCTURE ONLY -- inert stand-in for mpl_toolkits (ADI:20)."""
mplot3d = None
