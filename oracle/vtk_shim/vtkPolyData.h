/* stand-in for VTK's vtkPolyData.h: see vtk_standin.h (test infrastructure) */
#include "vtk_standin.h"
