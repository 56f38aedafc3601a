import sys
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parent.parent / "tests"))
import spice2_b200 as sp
from spice2_b200.samples import brunel
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
net, pops = brunel(plastic=True, mode=mode) if mode else brunel(plastic=True)
net.finalize()
net.step(300)
net.sync()
net.step(256)
net.sync()
net.close()
