import sys, time
t0=time.time()
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import termdaw_b200
from helpers import load_oracle
import test_pass_reuse as TR, test_playback as TPB
lib = termdaw_b200.load_library(); orc = load_oracle()
print("loaded %.1fs"%(time.time()-t0), flush=True)
TR.test_heads_only_reuse_of_sample_multi_and_lerp(lib, orc); print("heads-only OK %.1fs"%(time.time()-t0), flush=True)
TPB.check_play_equals_chunkwise_reference(lib, orc, "p3"); print("play OK %.1fs"%(time.time()-t0), flush=True)
TPB.check_stream_feed(lib, orc); print("feed OK %.1fs"%(time.time()-t0), flush=True)
TR.test_reuse_is_bit_identical(lib, orc, "p5"); print("reuse p5 OK %.1fs"%(time.time()-t0), flush=True)
