/* stand-in for VTK's vtkPointData.h: see vtk_standin.h (test infrastructure) */
#include "vtk_standin.h"
