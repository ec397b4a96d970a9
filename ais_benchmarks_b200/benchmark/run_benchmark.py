"""python -m ais_benchmarks_b200.benchmark.run_benchmark [benchmark.yaml methods.yaml config.yaml results.txt]
(the reference's ais_benchmarks/benchmark/run_benchmark.py with this package's CBenchmark)."""
import pathlib
import sys
import time
from os import sep

from ais_benchmarks_b200.benchmark import CBenchmark
from ais_benchmarks_b200.benchmark.CBenchmark import time_to_hms


def main(argv):
    def_path = str(pathlib.Path(__file__).parent.absolute()) + sep
    b_file, m_file, c_file = def_path + "def_benchmark.yaml", def_path + "def_methods.yaml", def_path + "def_config.yaml"
    out_file = "results" + sep + "results.txt"
    if len(argv) == 5:
        b_file, m_file, c_file, out_file = argv[1:5]
    t0 = time.time()
    CBenchmark().run(benchmark_file=b_file, methods_file=m_file, config_file=c_file, out_file=out_file)
    print("TOTAL EVALUATION TIME: %dh %dm %5.3fs" % time_to_hms(time.time() - t0))


if __name__ == "__main__":
    main(sys.argv)
