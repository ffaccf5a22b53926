/* stand-in for VTK's vtkCellData.h: see vtk_standin.h (test infrastructure) */
#include "vtk_standin.h"
