/* stand-in for VTK's vtkImageAlgorithm.h: see vtk_standin.h (test infrastructure) */
#include "vtk_standin.h"
