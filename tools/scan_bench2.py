import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.scan_bench import run
if __name__ == "__main__":
    run("f16", 1_000_000, 384, 10)
    run("f16", 4_000_000, 384, 10)
    run("i8", 12_500_000, 384, 10)
    run("i8", 12_500_000, 384, 100)
    run("b1", 32_000_000, 1024, 10)
    run("b1", 32_000_000, 1024, 100)
    run("b1", 16_000_000, 2048, 100)
    run("i8", 16_000_000, 128, 100)
    run("f16", 4_000_000, 384, 100)
