"""Host mirror of lv_slam::InformationMatrixCalculator (/root/reference/include/global_graph/information_matrix_calculator.hpp:11-60,
src/global_graph/information_matrix_calculator.cpp:8-87): the edge information matrix the pose-graph nodelet attaches to odometry
and loop edges (global_graph_nodelet.cpp:298, 697).  The nearest-neighbour fitness score runs on the GPU through the C-ABI
(lvs_ndt_fitness_score); the weighting is a handful of scalar operations and stays on the host like in the reference."""
import math

import numpy as np

from . import _capi as C
from .ndt import NormalDistributionsTransform

DBL_MAX = float(np.finfo(np.float64).max)


class InformationMatrixCalculator:
    """Parameter names of InformationMatrixCalculator; defaults of the constructor the nodelet uses,
    InformationMatrixCalculator(ros::NodeHandle&) (src/global_graph/information_matrix_calculator.cpp:8-21) - fitness_score_thresh is
    0.5 there (2.5 only in the unused load() template, information_matrix_calculator.hpp:32).  launch/dlo_lfa_ggo_kitti.launch:107
    sets the node's fitness_score_thresh to 2.0, which this object reads as well: pass fitness_score_thresh=2.0 to mirror that launch file."""

    def __init__(self, use_const_inf_matrix=False, const_stddev_x=0.5, const_stddev_q=0.1, var_gain_a=20.0, min_stddev_x=0.1,
                 max_stddev_x=5.0, min_stddev_q=0.05, max_stddev_q=0.2, fitness_score_thresh=0.5, device=0):
        self.use_const_inf_matrix = use_const_inf_matrix
        self.const_stddev_x, self.const_stddev_q = const_stddev_x, const_stddev_q
        self.var_gain_a = var_gain_a
        self.min_stddev_x, self.max_stddev_x = min_stddev_x, max_stddev_x
        self.min_stddev_q, self.max_stddev_q = min_stddev_q, max_stddev_q
        self.fitness_score_thresh = fitness_score_thresh
        self._device = device
        self._reg = None

    @staticmethod
    def weight(a, max_x, min_y, max_y, x):
        """information_matrix_calculator.hpp:40-44"""
        y = (1.0 - math.exp(-a * x)) / (1.0 - math.exp(-a * max_x))
        return min_y + (max_y - min_y) * y

    def calc_fitness_score(self, cloud1, cloud2, relpose, max_range=DBL_MAX):
        """Mean squared nearest-neighbour distance of relpose * cloud2 to cloud1 (information_matrix_calculator.cpp:53-87)."""
        if self._reg is None:
            self._reg = NormalDistributionsTransform(C.LVS_NDT_OMP, device=self._device)
        self._reg.setInputTarget(cloud1)
        self._reg.setInputSource(cloud2)
        return self._reg.getFitnessScore(max_range, T=np.asarray(relpose, dtype=np.float64).astype(np.float32))

    def information_from_fitness(self, fitness_score):
        """The part of calc_information_matrix after the fitness score (information_matrix_calculator.cpp:39-50); w_x and w_q are
        narrowed to float like the reference's `float w_x = weight(...)`."""
        w_x = np.float32(self.weight(self.var_gain_a, self.fitness_score_thresh, self.min_stddev_x ** 2, self.max_stddev_x ** 2, fitness_score))
        w_q = np.float32(self.weight(self.var_gain_a, self.fitness_score_thresh, self.min_stddev_q ** 2, self.max_stddev_q ** 2, fitness_score))
        inf = np.eye(6)
        inf[:3, :3] /= float(w_x)
        inf[3:, 3:] /= float(w_q)
        return inf

    def calc_information_matrix(self, cloud1, cloud2, relpose):
        if self.use_const_inf_matrix:
            inf = np.eye(6)
            inf[:3, :3] /= self.const_stddev_x
            inf[3:, 3:] /= self.const_stddev_q
            return inf
        return self.information_from_fitness(self.calc_fitness_score(cloud1, cloud2, relpose))
