"""bayesvg_b200 -- B200-native engine for bayesVG's SVG hot path (mean-field ADVI on the
hierarchical approximate GP of inst/approxGP*.stan).  Host-side mirror of the R interface in
`api`, C-ABI binding in `_lib`/`engine`, CUDA sources in `csrc/`."""
from .api import (SpatialData, VariationalFit, classifySVGs, clusterSVGsBayes, extractModel, findSpatiallyVariableFeaturesBayes, getBayesianGeneStats,  # noqa: F401
                  kmeans_centers, length_scale_kmeans, scale_coords)
from .engine import Fit, basis_build, comm_unique_id, default_config  # noqa: F401
