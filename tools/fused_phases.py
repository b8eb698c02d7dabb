"""Phase breakdown of the fused OPNet forward only (phases build)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "lstm_phases.py")).read()
head, tail = src.split("# fused OPNet forward")
exec(head.split("for H in (256, 512):")[0] + "\n" + tail)
