/* stand-in for VTK's vtkPoints.h: see vtk_standin.h (test infrastructure) */
#include "vtk_standin.h"
